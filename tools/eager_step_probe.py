"""Eager training steps on a PyG-layout batch (ncu target).  python tools/eager_step_probe.py [graphs] [steps] [pad_to_128]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-epc-saft_b200")):
    sys.path.insert(0, p)
import torch
from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
from gnnepcsaft_b200.train.engine import TrainEngine
from gnnepcsaft_b200.train.utils import create_model, default_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda")
cfg = default_config()
cfg.model = "PNA"
packed = make_packed_graphs(B, seed=1)
batch = batch_from_packed(packed).to(dev)
torch.manual_seed(0)
model = create_model(cfg, packed.in_degree_histogram()).to(dev)
eng = TrainEngine.from_config(model, cfg)
for _ in range(S):
    loss = eng.step(batch)
torch.cuda.synchronize()
print("nodes", batch.x.shape[0], "loss", float(loss))
