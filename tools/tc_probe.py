"""Per-role cycle breakdown of the tensor-core GEMM kernels (debug aid, run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-epc-saft_b200"))
import torch
from gnnepcsaft_b200 import _lib

dev = "cuda"
lib = _lib.load()
import ctypes
lib.gnn_tc_set_profile_buffer.argtypes = [ctypes.c_void_p]
prof = torch.zeros(16, dtype=torch.int64, device=dev)
ws = torch.zeros(4 << 20, dtype=torch.uint8, device=dev)
st = lambda: _lib.stream_ptr()
PREC = int(os.environ.get("TC_PREC", "0"))      # 0 = TF32x3, 1 = tf32 + bf16 cross terms, 2 = bf16


def report(name, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    lib.gnn_tc_set_profile_buffer(prof.data_ptr())
    prof.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    lib.gnn_tc_set_profile_buffer(None)
    p = prof.tolist()
    us = e0.elapsed_time(e1) * 1e3 / reps
    kb, mk, it, ct = max(p[3], 1), max(p[5], 1), max(p[8], 1), max(p[10], 1)
    print(f"{name:28s} {us:8.1f} us/call | per k-block (cycles): wait_stage {p[0]/kb:8.0f} produce {p[1]/kb:8.0f} "
          f"arrive {p[2]/kb:6.0f} | mma wait {p[4]/mk:8.0f} | per item: epi wait {p[6]/it:9.0f} epi work {p[7]/it:8.0f} "
          f"(tmem ld {p[11]/it:6.0f}) | CTA cycles {p[9]/ct:9.0f} ({p[9]/ct/1.9e3:.1f} us) kblocks/CTA {p[5]/ct:.1f} | first item: "
          f"start {p[12]/ct:7.0f} mma issued {p[13]/ct:8.0f} acc ready {p[14]/ct:8.0f}")


n = int(os.environ.get("TC_ROWS", "102263"))
x = torch.randn(n, 64, device=dev)
W = torch.randn(64, 64, device=dev) * 0.1
b = torch.randn(64, device=dev)
Y = torch.empty(n, 64, device=dev)
report("linear_fwd 64x64", lambda: _lib.call("gnn_tc_linear_fwd", x, 64, W, 64, b, n, 64, 64, 0, Y, 64, PREC, ws, ws.numel(), st()))
report("linear_fwd 64x64 bf16", lambda: _lib.call("gnn_tc_linear_fwd", x, 64, W, 64, b, n, 64, 64, 0, Y, 64, 2, ws, ws.numel(), st()))
dW = torch.zeros(64, 64, device=dev)
db = torch.zeros(64, device=dev)
report("linear_wgrad 64x64", lambda: _lib.call("gnn_tc_linear_wgrad", Y, 64, x, 64, n, 64, 64, dW, 64, db, PREC, st()))
F, Fo, T = 64, 32, 2
A = torch.randn(n, T * 4 * F, device=dev)
scal = torch.rand(n, 2, device=dev) + 0.5
wts = Fo * 13 * F + Fo
wpost = torch.randn(T * wts, device=dev) * 0.1
O = torch.empty(n, T * Fo, device=dev)
report("post_fwd", lambda: _lib.call("gnn_tc_pna_post_fwd", x, A, scal, wpost, wts, wpost[Fo * 13 * F:], wts, n, F, Fo, T, 0, O, PREC, ws, ws.numel(), st()))
gA = torch.empty(n, T * 4 * F, device=dev)
report("post_dgrad gA", lambda: _lib.call("gnn_tc_pna_post_dgrad", O, scal, wpost, wts, n, F, Fo, T, None, gA, PREC, ws, ws.numel(), st()))
dwp = torch.zeros(T * wts, device=dev)
report("post_wgrad (2 launches)", lambda: _lib.call("gnn_tc_pna_post_wgrad", O, x, A, scal, n, F, Fo, T, dwp, wts, dwp[Fo * 13 * F:], wts, PREC, st()))
C = T * F
wpre_ts = F * 3 * F + F                      # per-tower pre-NN weight [F, 3F] + bias [F]
wpre = torch.randn(T * wpre_ts, device=dev) * 0.1
P = torch.empty(n, 2 * C, device=dev)
report("proj_fwd", lambda: _lib.call("gnn_tc_pna_proj_fwd", x, wpre, wpre_ts, wpre[F * 3 * F:], wpre_ts, n, F, T, P, PREC, ws, ws.numel(), st()))
dx = torch.empty(n, F, device=dev)
report("dx_dgrad", lambda: _lib.call("gnn_tc_pna_dx_dgrad", P, O, wpre, wpre_ts, wpost, wts, n, F, Fo, T, 0, dx, PREC, ws, ws.numel(), st()))
dwpre = torch.zeros(T * wpre_ts, device=dev)
report("proj_wgrad", lambda: _lib.call("gnn_tc_pna_proj_wgrad", P, x, n, F, T, dwpre, wpre_ts, dwpre[F * 3 * F:], wpre_ts, PREC, st()))
