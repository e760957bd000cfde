"""Runs the fused message+aggregate kernels alone on a BASELINE config-2 sized batch (for ncu / timing).

  python tools/agg_probe.py [reps]         env: GNN_B200_AGG_FWD (kernel variant), GNN_B200_AGG_CTAS
Two cache states per kernel: "cold" (L2 flushed by a 256 MB fill before every launch) and "step-like"
(the projection tensor written by a device copy right before the launch, as the projection GEMM does).
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-epc-saft_b200"))
import torch
from gnnepcsaft_b200 import _lib
from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
from gnnepcsaft_b200.train.models import GraphStructure

dev = "cuda"
packed = make_packed_graphs(4096, seed=1)
b = batch_from_packed(packed).to(dev)
n = b.x.shape[0]
g = GraphStructure(b.edge_index, b.edge_attr, b.batch, n, 4096, True)
T, F = 2, 64
C = T * F
P = torch.randn(n, 2 * C, device=dev)
Ps = P[:, C:].contiguous()
Ps_src = Ps.clone()
tab = torch.randn(60, C, device=dev)
A = torch.empty(n, 4 * C, device=dev)
gA = torch.randn(n, 4 * C, device=dev)
dq = torch.empty(g.ep, C, device=dev)
dPs = torch.empty(n, C, device=dev)
dtab = torch.zeros(60, C, device=dev)
nb = _lib.query("gnn_pna_aggregate_bwd_workspace_bytes", T, F)
zws = torch.zeros(nb, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ep = g.ep
alg = 4 * n * 64 + 16 * n * T * F + 5 * ep + 4 * (n + 1)
st = _lib.stream_ptr


def timed(name, fn, pre):
    times = []
    for i in range(reps):
        pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(times)[len(times) // 2]
    print(f"{name:44s} median {t:6.1f} us   algorithmic {alg/1e6:.1f} MB -> {alg/t/1e3:5.0f} GB/s = {alg/t/1e3/6535.7:.3f} of measured peak")
    return t


def cold():
    flush.zero_()


def steplike():
    flush.zero_()
    Ps.copy_(Ps_src)          # Psrc freshly written (dirty in L2), as after the projection GEMM


print(f"nodes {n} slots {ep} variant {os.environ.get('GNN_B200_AGG_FWD', '0')} ctas {os.environ.get('GNN_B200_AGG_CTAS', '-')}")
timed("aggregate_fwd [Pdst|Psrc] cold", lambda: _lib.call("gnn_pna_aggregate_fwd", P, tab, g.rowptr, g.col_src, g.col_code, n, T, F, A, st()), cold)
timed("aggregate_src_fwd cold", lambda: _lib.call("gnn_pna_aggregate_src_fwd", Ps, tab, g.rowptr, g.col_src, g.col_code, n, T, F, A, st()), cold)
timed("aggregate_src_fwd step-like", lambda: _lib.call("gnn_pna_aggregate_src_fwd", Ps, tab, g.rowptr, g.col_src, g.col_code, n, T, F, A, st()), steplike)
timed("aggregate_src_bwd cold", lambda: _lib.call("gnn_pna_aggregate_src_bwd", gA, Ps, tab, A, g.rowptr, g.col_src, g.col_code, n, T, F, dPs, dq, dtab, zws, nb, st()), cold)
timed("aggregate_src_bwd atomic scatter (default) cold", lambda: _lib.call("gnn_pna_aggregate_src_bwd", gA, Ps, tab, A, g.rowptr, g.col_src, g.col_code, n, T, F, dPs, None, dtab, zws, nb, st()), cold)
timed("slot_sum_by_source cold", lambda: _lib.call("gnn_slot_sum_by_source", dq, g.srcptr, g.src_slot, n, C, dPs, C, st()), cold)
