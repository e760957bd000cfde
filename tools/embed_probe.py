"""AtomEncoder backward alone: python tools/embed_probe.py [nodes] [hidden]  (CUDA-event time per call, L2-warm)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-epc-saft_b200")):
    sys.path.insert(0, p)
import torch
from gnnepcsaft_b200 import _lib
from gnnepcsaft_b200.data.synth import make_packed_graphs

n_target = int(sys.argv[1]) if len(sys.argv) > 1 else 102400
H = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda")
packed = make_packed_graphs(max(n_target // 24, 1), seed=3)
x = packed.x[:n_target].to(torch.uint8).to(dev)
n = x.shape[0]
off = torch.tensor([0, 119, 124, 136, 148, 158, 164, 170, 172], dtype=torch.int32, device=dev)
dout = torch.randn(n, H, device=dev)
dt = torch.zeros(174, H, device=dev)
for name, xx in (("gnn_atom_embed_u8_bwd", x), ("gnn_atom_embed_bwd", x.long())):
    for _ in range(3):
        _lib.call(name, xx, n, dout, off, 174, H, dt, _lib.stream_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        _lib.call(name, xx, n, dout, off, 174, H, dt, _lib.stream_ptr())
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: n={n} H={H}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us/call")
