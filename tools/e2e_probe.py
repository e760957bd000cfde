"""Where does the end-to-end leg lose time against the device-resident step?  Variants of the bench.py e2e loop."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gnn-epc-saft_b200"))
import torch
import bench
from gnnepcsaft_b200.data.loader import DevicePrefetcher
from gnnepcsaft_b200.train.engine import TrainEngine
from gnnepcsaft_b200.train.utils import create_model

dev = torch.device("cuda")
packed, cfg, host_batches = bench.build_problem(seed=1)
torch.manual_seed(0)
model = create_model(cfg, packed.in_degree_histogram()).to(dev)
engine = TrainEngine.from_config(model, cfg)
dev_batches = [b.to(dev) for b in host_batches]
pinned = [b.pin_memory() for b in host_batches]
K = 20
slots = torch.empty(K + 3, dtype=torch.float32, pin_memory=True)


def feed(count):
    for i in range(count):
        yield pinned[i % 4]


def run(name, h2d, readback, depth=2):
    def body(count):
        pending = None
        it = DevicePrefetcher(feed(count), dev, depth=depth) if h2d else (dev_batches[i % 4] for i in range(count))
        for k, db in enumerate(it):
            loss = engine.step(db)
            if readback:
                slots[k:k + 1].copy_(loss, non_blocking=True)
                ev = torch.cuda.Event(); ev.record()
                if pending is not None:
                    pending.synchronize()
                pending = ev
        if pending is not None:
            pending.synchronize()
    body(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(); body(K); e1.record()
    torch.cuda.synchronize()
    print(f"{name:46s} {e0.elapsed_time(e1) / K:6.3f} ms/step (wall {1e3 * (time.perf_counter() - t0) / K:6.3f})")


run("device batches, no readback", False, False)
run("device batches, loss readback one step late", False, True)
run("prefetched host batches, no readback", True, False)
run("prefetched host batches + readback (bench e2e)", True, True)
run("prefetched host batches + readback, depth 3", True, True, depth=3)
