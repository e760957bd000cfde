"""Store-resident training step replayed from its captured CUDA graph (target of the ncu launch list / --set full
captures of round 2).  python tools/graph_step_probe.py [graphs_per_batch] [replays]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-epc-saft_b200")):
    sys.path.insert(0, p)
import torch
from gnnepcsaft_b200.data.compact import CompactGraphStore
from gnnepcsaft_b200.data.synth import make_packed_graphs
from gnnepcsaft_b200.train.engine import TrainEngine
from gnnepcsaft_b200.train.utils import create_model, default_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
R = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda")
cfg = default_config()
cfg.model = "PNA"
packed = make_packed_graphs(4 * B, seed=1)
store = CompactGraphStore(packed, dev, keep_host_mirror=False)
torch.manual_seed(0)
model = create_model(cfg, packed.in_degree_histogram()).to(dev)
eng = TrainEngine.from_config(model, cfg)
ids = torch.arange(B)
for _ in range(2):                 # eager warm-up step, then capture + first replay
    eng.step_from_store(store, ids)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(R):
    loss = eng.step_from_store(store, ids)
e1.record()
torch.cuda.synchronize()
print(f"{B} graphs: {e0.elapsed_time(e1) / R:.3f} ms/step over {R} replays, loss {float(loss):.5f}, "
      f"replays {eng.graph_replays}")
