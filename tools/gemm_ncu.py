"""One launch each of the three PNA post-stage tensor-core GEMMs at the 4096-graph batch size (for `ncu --set full`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-epc-saft_b200"))
import torch
from gnnepcsaft_b200 import _lib

dev = "cuda"
n, F, Fo, T = 102263, 64, 32, 2
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(n, F, device=dev, generator=g)
A = torch.randn(n, T * 4 * F, device=dev, generator=g)
scal = torch.rand(n, 2, device=dev, generator=g) + 0.5
wts = Fo * 13 * F + Fo
wpost = torch.randn(T * wts, device=dev, generator=g) * 0.1
Vx = torch.randn(T, 3, Fo, F, device=dev, generator=g) * 0.1
O = torch.empty(n, T * Fo, device=dev)
dO = torch.randn(n, T * Fo, device=dev, generator=g)
gA = torch.empty(n, T * 4 * F, device=dev)
dw = torch.zeros(T * wts, device=dev)
dVx = torch.zeros(T, 3, Fo, F, device=dev)
ws = torch.zeros(32 << 20, dtype=torch.uint8, device=dev)
st = _lib.stream_ptr()
for _ in range(2):
    _lib.call("gnn_tc_pna_postf_fwd", x, A, scal, Vx, wpost, wts, wpost[Fo * 13 * F:], wts, n, F, Fo, T, 0, O, 0, ws, ws.numel(), st)
    _lib.call("gnn_tc_pna_post_dgrad", dO, scal, wpost, wts, n, F, Fo, T, None, gA, 0, ws, ws.numel(), st)
    _lib.call("gnn_tc_pna_postf_wgrad", dO, x, A, scal, n, F, Fo, T, dw, wts, dw[Fo * 13 * F:], wts, dVx, 0, st)
torch.cuda.synchronize()
print("ok")
