"""Runs the fused message+aggregate forward kernel alone on a BASELINE config-2 sized batch (for ncu / timing)."""
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
tab = torch.randn(60, C, device=dev)
A = torch.empty(n, 4 * C, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
times = []
for i in range(reps):
    flush.zero_()                                   # evict L2 between launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call("gnn_pna_aggregate_fwd", P, tab, g.rowptr, g.col_src, g.col_code, n, T, F, A, _lib.stream_ptr())
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) * 1e3)
ep = g.ep
alg = 4 * n * 64 + 16 * n * T * F + 5 * ep + 4 * (n + 1)
t = sorted(times)[len(times) // 2]
print(f"nodes {n} slots {ep}: median {t:.1f} us  algorithmic {alg/1e6:.1f} MB -> {alg/t/1e3:.0f} GB/s ({alg/t/1e3/6535.7:.3f} of measured peak); "
      f"actual traffic {(n*2*C*4 + n*4*C*4)/1e6:.0f} MB -> {(n*2*C*4 + n*4*C*4)/t/1e3:.0f} GB/s")
