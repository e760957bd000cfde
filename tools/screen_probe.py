"""Screening throughput (engine.screen) against the micro-batch node budget (run on the GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-epc-saft_b200")):
    sys.path.insert(0, p)
import torch
from gnnepcsaft_b200.data.compact import CompactGraphStore
from gnnepcsaft_b200.data.synth import make_packed_graphs
from gnnepcsaft_b200.train.engine import TrainEngine
from gnnepcsaft_b200.train.utils import create_model, default_config

dev = torch.device("cuda")
cfg = default_config()
cfg.model = "PNA"
packed = make_packed_graphs(32768, seed=77)
store = CompactGraphStore(packed, dev, keep_host_mirror=False)
torch.manual_seed(0)
model = create_model(cfg, packed.in_degree_histogram()).to(dev)
eng = TrainEngine.from_config(model, cfg)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for max_nodes in [int(a) for a in (sys.argv[1:] or ["65536", "131072", "262144", "524288"])]:
    for _ in range(2):
        _, nmb, nodes = eng.screen(store, max_nodes=max_nodes, out=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    eng.screen(store, max_nodes=max_nodes, out=False, passes=4)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    print(f"max_nodes {max_nodes:7d}: {nmb:3d} micro-batches, real nodes {sum(nodes)}, padded {nmb * max_nodes}, "
          f"{ms:7.2f} ms/pass = {32768 / ms * 1e3 / 1e6:.2f} M graphs/s (host wall {1e3 * (time.perf_counter() - t0) / 4:.2f} ms/pass)")
