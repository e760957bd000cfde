"""Times gnn_tc_pna_postf_agg_fwd against the two-kernel path on a 4096-graph batch (run on the GPU box; also the
target of the ncu captures of the fused kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-epc-saft_b200")):
    sys.path.insert(0, p)
import torch
from gnnepcsaft_b200 import _lib
if os.environ.get("PROF"):      # per-role cycle counters: python gnn-epc-saft_b200/csrc/build.py --profile
    _lib.LIB_PATH = _lib.LIB_PATH.replace(".so", "_prof.so")
from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
from gnnepcsaft_b200.train.models import GraphStructure

dev = torch.device("cuda")
graphs = int(os.environ.get("GRAPHS", "4096"))
packed = make_packed_graphs(graphs, seed=1)
b = batch_from_packed(packed).to(dev)
n = b.x.shape[0]
g = GraphStructure(b.edge_index, b.edge_attr, b.batch, n, graphs, True)
F, T, Fo = 64, 2, 32
C = T * F
gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(n, F, device=dev, generator=gen)
P = torch.randn(n, C, device=dev, generator=gen)
tab = torch.randn(60, C, device=dev, generator=gen) * 0.3
scal = torch.rand(n, 2, device=dev, generator=gen) + 0.5
post_ts = Fo * 13 * F + Fo + 4
post = torch.randn(T * post_ts, device=dev, generator=gen) * 0.1
Vx = torch.randn(T, 3, Fo, F, device=dev, generator=gen) * 0.1
boff = Fo * 13 * F
ws = torch.zeros(32 << 20, dtype=torch.uint8, device=dev)
O = torch.zeros(n, T * Fo, device=dev)
A = torch.zeros(n, T * 4 * F, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = _lib.stream_ptr


def fused(store):
    _lib.call("gnn_tc_pna_postf_agg_fwd", x, P, tab, g.rowptr, g.col_src, g.col_code, scal, Vx, post, post_ts, post[boff:],
              post_ts, n, F, Fo, T, 0, O, A if store else None, 0, ws, ws.numel(), st())


def two_kernel():
    _lib.call("gnn_pna_aggregate_src_fwd", P, tab, g.rowptr, g.col_src, g.col_code, n, T, F, A, st())
    _lib.call("gnn_tc_pna_postf_fwd", x, A, scal, Vx, post, post_ts, post[boff:], post_ts, n, F, Fo, T, 0, O, 0, ws, ws.numel(), st())


def timed(name, fn, reps=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    print(f"{name:40s} median {ts[len(ts) // 2]:8.1f} us  min {ts[0]:8.1f} us  (n = {n})")


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "fused"):
    timed("fused agg+post fwd (no A store)", lambda: fused(False))
    timed("fused agg+post fwd (A stored)", lambda: fused(True))
if which in ("all", "two"):
    timed("aggregate_src_fwd + postf_fwd", two_kernel)

if os.environ.get("PROF"):
    import ctypes
    lib = _lib.load()
    lib.gnn_tc_set_profile_buffer.argtypes = [ctypes.c_void_p]
    prof = torch.zeros(16, dtype=torch.int64, device=dev)
    for name, fn in (("fused", lambda: fused(False)), ("fused+A", lambda: fused(True)),
                     ("postf only", lambda: _lib.call("gnn_tc_pna_postf_fwd", x, A, scal, Vx, post, post_ts, post[boff:], post_ts, n, F, Fo, T, 0, O, 0, ws, ws.numel(), st()))):
        fn(); torch.cuda.synchronize()
        lib.gnn_tc_set_profile_buffer(prof.data_ptr()); prof.zero_()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        lib.gnn_tc_set_profile_buffer(None)
        p = prof.tolist()
        kb, mk, it, ct = max(p[3], 1), max(p[5], 1), max(p[8], 1), max(p[10], 1)
        print(f"{name:12s} per k-block (cycles): wait_stage(+gather) {p[0]/kb:8.0f} store {p[1]/kb:8.0f} arrive {p[2]/kb:6.0f} | "
              f"gather+reduce per agg k-block {p[15]/max(kb*0.8,1):8.0f} | mma wait {p[4]/mk:8.0f} | per item: epi wait {p[6]/it:9.0f} "
              f"epi work {p[7]/it:8.0f} | CTA cycles {p[9]/ct:9.0f} kblocks/CTA {p[5]/ct:.1f}")
