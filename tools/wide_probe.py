"""BASELINE config 5 sanity / throughput probe: wide model (hidden 256, 6 conv layers, post_layers 2, no self-loops;
reference configs_parallel.py:43-51) on skewed graph sizes (5-150 atoms), training step and eval forward.

  python tools/wide_probe.py [graphs_per_batch]      env GNN_B200_MATMUL = tf32x3 | bf16
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-epc-saft_b200"))
import torch
from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
from gnnepcsaft_b200.train.engine import TrainEngine
from gnnepcsaft_b200.train.utils import create_model, default_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda")
cfg = default_config()
cfg.model = "PNA"
cfg.hidden_dim, cfg.propagation_depth, cfg.post_layers, cfg.add_self_loops = 256, 6, 2, False
packed = make_packed_graphs(2 * B, seed=4, skewed=True)
batches = [batch_from_packed(packed, torch.arange(i * B, (i + 1) * B)).to(dev) for i in range(2)]
torch.manual_seed(0)
model = create_model(cfg, packed.in_degree_histogram()).to(dev)
print("params", sum(p.numel() for p in model.parameters()), "nodes/batch", int(batches[0].x.shape[0]),
      "matmul", model.matmul_precision, "folded", model._fold_dst(model._matmul_mode()))
eng = TrainEngine.from_config(model, cfg)
for i in range(3):
    loss = eng.step(batches[i % 2])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 10
for i in range(K):
    loss = eng.step(batches[i % 2])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"train: {ms:.2f} ms/step = {B / ms * 1e3:,.0f} graphs/s, loss {float(loss):.4f}")
for i in range(2):
    eng.predict(batches[i])
torch.cuda.synchronize()
e0.record()
for i in range(K):
    out = eng.predict(batches[i % 2])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"eval forward: {ms:.2f} ms/batch = {B / ms * 1e3:,.0f} graphs/s, finite {bool(torch.isfinite(out).all())}")
