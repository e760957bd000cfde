"""gnn_tc_linear_fwd (K = N = 64, the `lin` layer) against the number of rows: fixed cost vs per-tile cost."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-epc-saft_b200"))
import torch
from gnnepcsaft_b200 import _lib
dev = "cuda"
W = torch.randn(64, 64, device=dev) * 0.1
b = torch.randn(64, device=dev)
ws = torch.empty(32 << 20, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = _lib.stream_ptr
for M in (12800, 102400, 409600, 1638400):
    X = torch.randn(M, 64, device=dev)
    Y = torch.empty(M, 64, device=dev)
    for prepacked in (False, True):
        nb = -ws.numel() if prepacked else ws.numel()
        _lib.call("gnn_tc_linear_fwd", X, 64, W, 64, b, M, 64, 64, 0, Y, 64, 0, ws, ws.numel(), st())
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.call("gnn_tc_linear_fwd", X, 64, W, 64, b, M, 64, 64, 0, Y, 64, 0, ws, nb, st())
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        t = ts[len(ts) // 2]
        print(f"M {M:8d} prepacked {int(prepacked)}: {t:7.1f} us  = {t / (M / 128 / 148):6.2f} us per tile-round, {2 * M * 256 / t / 1e3:6.0f} GB/s")
