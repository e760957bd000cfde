"""Host-side cost of one predict_from_store / step_from_store call (graph replay path), against the GPU time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-epc-saft_b200")):
    sys.path.insert(0, p)
import torch
from gnnepcsaft_b200.data.compact import CompactGraphStore, bucket
from gnnepcsaft_b200.data.synth import make_packed_graphs
from gnnepcsaft_b200.train.engine import TrainEngine
from gnnepcsaft_b200.train.utils import create_model, default_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda")
cfg = default_config(); cfg.model = "PNA"
packed = make_packed_graphs(8 * B, seed=1)
store = CompactGraphStore(packed, dev, keep_host_mirror=False)
torch.manual_seed(0)
model = create_model(cfg, packed.in_degree_histogram()).to(dev)
eng = TrainEngine.from_config(model, cfg)
gen = torch.Generator().manual_seed(0)
batches = [torch.randperm(8 * B, generator=gen)[:B] for _ in range(40)]
dev_ids = [b.to(dev) for b in batches]
for fn_name in ("predict_from_store", "step_from_store"):
    fn = getattr(eng, fn_name)
    for ids, d in zip(batches, dev_ids):          # warm-up: capture every bucket
        fn(store, ids, ids_device=d)
        fn(store, ids, ids_device=d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for ids, d in zip(batches, dev_ids):
        fn(store, ids, ids_device=d)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{fn_name} {B} graphs: host enqueue {1e3 * (t1 - t0) / len(batches):.3f} ms/call, GPU {e0.elapsed_time(e1) / len(batches):.3f} ms/call")
# pieces of the host path
ids, d = batches[0], dev_ids[0]
def t(f, n=200):
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    return 1e6 * (time.perf_counter() - t0) / n
print(f"store.sizes {t(lambda: store.sizes(ids)):.1f} us, bucket {t(lambda: bucket(*store.sizes(ids))):.1f} us, "
      f"plan token {t(lambda: model._infer_plan(dev)):.1f} us, arena() {t(lambda: model.arena()):.1f} us")
