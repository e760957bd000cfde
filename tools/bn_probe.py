"""Times the BatchNorm statistics / backward entry points alone at the 4096-graph batch size (L2 warm and cold)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-epc-saft_b200"))
import torch
from gnnepcsaft_b200 import _lib

dev = "cuda"
rows, C = 102263, 64
Y = torch.randn(rows, C, device=dev)
dO = torch.randn(rows, C, device=dev)
dY = torch.empty_like(Y)
out = torch.empty_like(Y)
ga, be = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
nbt = torch.zeros(1, dtype=torch.int64, device=dev)
stats = torch.empty(2 * C, device=dev)
dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
wsb = _lib.query("gnn_bn_workspace_bytes", rows, C)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s = _lib.stream_ptr()
calls = {
    "bn_stats": lambda: _lib.call("gnn_bn_stats", Y, rows, C, 1, 1e-5, 0.1, rm, rv, nbt, stats, None, ws, wsb, s),
    "bn_act_fwd": lambda: _lib.call("gnn_bn_act_fwd", Y, stats, ga, be, dO, rows, C, 1, out, s),
    "bn_act_bwd": lambda: _lib.call("gnn_bn_act_bwd", dO, Y, stats, ga, be, rows, C, 1, 1, dY, dg, db, None, ws, wsb, s),
}
for name, fn in calls.items():
    for cold in (0, 1):
        ts = []
        for i in range(12):
            if cold:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"{name:12s} {'cold' if cold else 'warm'}: median {sorted(ts)[len(ts)//2]:.1f} us")
