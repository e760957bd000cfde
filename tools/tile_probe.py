"""Tile inference kernel (gnn_tc_pna_stack_infer) against the per-layer path: parity and timing (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-epc-saft_b200")):
    sys.path.insert(0, p)
import torch
from gnnepcsaft_b200.data.compact import CompactGraphStore
from gnnepcsaft_b200.data.synth import make_packed_graphs
from gnnepcsaft_b200.train import models as M
from gnnepcsaft_b200.train.engine import TrainEngine
from tests.helpers import cfg, make_pair

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda")
c = cfg()
packed = make_packed_graphs(max(B, 64), seed=3, num_para=c.num_para)
_, _, m = make_pair(c, packed.in_degree_histogram(), dtype64=False)
m.to(dev)
eng = TrainEngine.from_config(m, c, distributed=False)
eng.use_cuda_graphs = False
store = CompactGraphStore(packed, dev)
print("tileable", store.tileable, flush=True)
ids = torch.arange(B)
res = {}
for tiles in (False, True):
    M.TILE_INFER = tiles
    y = eng.predict_from_store(store, ids).clone()
    torch.cuda.synchronize()
    res[tiles] = y
    print("tiles", tiles, "finite", bool(torch.isfinite(y).all()), flush=True)
err = float((res[True] - res[False]).abs().max() / res[False].abs().max())
print(f"tile kernel vs per-layer kernels: max-normalised difference {err:.3e}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for tiles in (False, True):
    M.TILE_INFER = tiles
    for _ in range(3):
        eng.predict_from_store(store, ids)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        eng.predict_from_store(store, ids)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"tiles={tiles}: {ms:.3f} ms per {B}-graph batch = {B / ms * 1e3 / 1e6:.2f} M graphs/s (eager)")

# captured-graph replay (the way bench.py / screen() run it)
for tiles in (False, True):
    M.TILE_INFER = tiles
    eng2 = TrainEngine.from_config(m, c, distributed=False)
    for _ in range(3):
        eng2.predict_from_store(store, ids)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(30):
        eng2.predict_from_store(store, ids)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print(f"tiles={tiles}: {ms:.3f} ms per {B}-graph batch = {B / ms * 1e3 / 1e6:.2f} M graphs/s (graph replay, {eng2.graph_replays} replays)")
