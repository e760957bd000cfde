"""Where does the end-to-end leg lose time against the store-resident step?  Variants of the bench.py e2e loop
(compact pinned host batches -> DevicePrefetcher -> step_from_host -> loss readback)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gnn-epc-saft_b200"))
import torch
import bench
from gnnepcsaft_b200.data.compact import CompactGraphStore
from gnnepcsaft_b200.data.loader import DevicePrefetcher
from gnnepcsaft_b200.train.engine import TrainEngine
from gnnepcsaft_b200.train.utils import create_model

dev = torch.device("cuda")
packed, cfg, _ = bench.build_problem(seed=1)
torch.manual_seed(0)
model = create_model(cfg, packed.in_degree_histogram()).to(dev)
engine = TrainEngine.from_config(model, cfg)
store = CompactGraphStore(packed, dev)
gen = torch.Generator().manual_seed(0)
ids_list = [torch.randperm(store.num_graphs, generator=gen)[:4096] for _ in range(8)]
host = [store.host_batch(i) for i in ids_list]
devb = [h.to(dev) for h in host]
print("buckets", sorted({(h.pad_nodes, h.pad_slots) for h in host}))
K = 40
slots = torch.empty(K + 8, dtype=torch.float32, pin_memory=True)


def run(name, mode, readback, depth=2):
    def body(count):
        pending = None
        if mode == "store":
            it = ((ids_list[i % 8],) for i in range(count))
        elif mode == "device":
            it = ((devb[i % 8],) for i in range(count))
        else:
            it = ((b,) for b in DevicePrefetcher((host[i % 8] for i in range(count)), dev, depth=depth))
        if mode == "feed":
            it = ((l,) for l in engine.steps_from_host(store, (host[i % 8] for i in range(count)), depth=depth))
        for k, (b,) in enumerate(it):
            loss = b if mode == "feed" else (engine.step_from_store(store, b) if mode == "store" else engine.step_from_host(store, b))
            if readback:
                slots[k:k + 1].copy_(loss, non_blocking=True)
                ev = torch.cuda.Event(); ev.record()
                if pending is not None:
                    pending.synchronize()
                pending = ev
        if pending is not None:
            pending.synchronize()
    body(16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(); body(K); e1.record()
    torch.cuda.synchronize()
    print(f"{name:52s} {e0.elapsed_time(e1) / K:6.3f} ms/step (wall {1e3 * (time.perf_counter() - t0) / K:6.3f})", flush=True)


for rep in range(2):
    run("store-resident (device collate), no readback", "store", False)
    run("device compact batches (D2D into static), no readback", "device", False)
    run("device compact batches + loss readback", "device", True)
    run("prefetched host batches, no readback", "host", False)
    run("prefetched host batches + readback (bench e2e)", "host", True)
    run("prefetched host batches + readback, depth 3", "host", True, depth=3)
    run("steps_from_host (static staging) + readback", "feed", True)
