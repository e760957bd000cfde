"""Wide model (config 5) on compact store batches: eager, then graph replay (debug aid; run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-epc-saft_b200")):
    sys.path.insert(0, p)
import torch
from gnnepcsaft_b200.data.compact import CompactGraphStore
from gnnepcsaft_b200.data.synth import make_packed_graphs
from gnnepcsaft_b200.train.engine import TrainEngine
from gnnepcsaft_b200.train.utils import create_model, wide_config

dev = torch.device("cuda")
wcfg = wide_config()
wcfg.model = "PNA"
wpacked = make_packed_graphs(4096, seed=4, skewed=True)
wstore = CompactGraphStore(wpacked, dev, self_loops=wcfg.add_self_loops, keep_host_mirror=False)
gen = torch.Generator().manual_seed(0)
for mode in (sys.argv[1:] or ["bf16", "tf32x3"]):
    torch.manual_seed(0)
    m = create_model(wcfg, wpacked.in_degree_histogram()).to(dev)
    m.matmul_precision = mode
    eng = TrainEngine.from_config(m, wcfg)
    eng.use_cuda_graphs = os.environ.get("GRAPHS", "1") == "1"
    for it in range(6):
        ids = torch.randperm(4096, generator=gen)[:1024]
        loss = eng.step_from_store(wstore, ids)
        torch.cuda.synchronize()
        print(mode, it, "loss", float(loss), "replays", eng.graph_replays, flush=True)
    y = eng.predict_from_store(wstore, ids)
    torch.cuda.synchronize()
    print(mode, "predict finite", bool(torch.isfinite(y).all()), flush=True)

# no sync between steps, many buckets (the pattern of bench.py's config5 leg)
from gnnepcsaft_b200.data.compact import bucket
for mode in (sys.argv[1:] or ["bf16", "tf32x3"]):
    torch.manual_seed(0)
    m = create_model(wcfg, wpacked.in_degree_histogram()).to(dev)
    m.matmul_precision = mode
    eng = TrainEngine.from_config(m, wcfg)
    batches = [torch.randperm(4096, generator=gen)[:1024] for _ in range(10)]
    print("buckets", sorted({bucket(*wstore.sizes(i)) for i in batches}))
    for rep in range(3):
        for ids in batches:
            loss = eng.step_from_store(wstore, ids)
        torch.cuda.synchronize()
        print(mode, "rep", rep, "loss", float(loss), "replays", eng.graph_replays, "graphs", len(eng._buckets), flush=True)
