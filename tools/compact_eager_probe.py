"""Eager training steps on a compact store batch (ncu target).  python tools/compact_eager_probe.py graphs steps pad(0|1|N)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-epc-saft_b200")):
    sys.path.insert(0, p)
import torch
from gnnepcsaft_b200.data.compact import CompactGraphStore
from gnnepcsaft_b200.data.synth import make_packed_graphs
from gnnepcsaft_b200.train.engine import TrainEngine
from gnnepcsaft_b200.train.utils import create_model, default_config

B, S, PAD = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda")
cfg = default_config()
cfg.model = "PNA"
packed = make_packed_graphs(B, seed=1)
store = CompactGraphStore(packed, dev, keep_host_mirror=False)
torch.manual_seed(0)
model = create_model(cfg, packed.in_degree_histogram()).to(dev)
eng = TrainEngine.from_config(model, cfg)
ids = torch.arange(B)
n, s = store.sizes(ids)
pad = False if PAD == 0 else (True if PAD == 1 else (n + PAD, s + 64))
cb = store.batch(ids, pad=pad)
for _ in range(S):
    loss = eng.step(cb)
torch.cuda.synchronize()
print("nodes", n, "padded", cb.pad_nodes, "loss", float(loss))
