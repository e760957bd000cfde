"""Call-by-call comparison of an exact and a padded compact step (finds the first kernel whose real rows differ)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-epc-saft_b200")):
    sys.path.insert(0, p)
import torch
from gnnepcsaft_b200.data.compact import CompactGraphStore, bucket
from gnnepcsaft_b200.data.synth import make_packed_graphs
from gnnepcsaft_b200.train import models as M, engine as E
from gnnepcsaft_b200.train.engine import TrainEngine
from tests.helpers import cfg, make_pair

DEV = "cuda"
c = cfg()
packed = make_packed_graphs(300, seed=2, num_para=c.num_para)
deg = packed.in_degree_histogram()
ids = torch.randperm(300, generator=torch.Generator().manual_seed(0))[:256]
store = CompactGraphStore(packed, device=DEV, self_loops=c.add_self_loops)
n, s = store.sizes(ids)
orig = M._c
logs = {}
for name, pad in (("exact", False), ("padded", True)):
    _, _, m = make_pair(c, deg, dtype64=False)
    m.to(DEV)
    e = TrainEngine.from_config(m, c, distributed=False)
    e.use_cuda_graphs = False
    cb = store.batch(ids, pad=pad)
    log = []

    def rec(fn, *args, _log=log):
        orig(fn, *args)
        torch.cuda.synchronize()
        _log.append((fn, [a.detach().clone() if isinstance(a, torch.Tensor) and a.is_floating_point() else None for a in args]))
    M._c = rec
    E._c = rec
    e.step(cb)
    M._c = orig
    E._c = orig
    logs[name] = (log, cb.pad_nodes)
(la, na), (lb, nb_) = logs["exact"], logs["padded"]
assert len(la) == len(lb)
for i, ((fa, aa), (fb, ab)) in enumerate(zip(la, lb)):
    assert fa == fb
    worst = 0.0
    what = ""
    for j, (ta, tb) in enumerate(zip(aa, ab)):
        if ta is None or tb is None:
            continue
        if ta.numel() == tb.numel():
            a, b = ta.reshape(-1), tb.reshape(-1)
        elif ta.numel() % na == 0 and tb.numel() % nb_ == 0 and ta.numel() // na == tb.numel() // nb_:
            w = ta.numel() // na
            a, b = ta.reshape(na, w)[:n].reshape(-1), tb.reshape(nb_, w)[:n].reshape(-1)
        elif ta.numel() % 257 == 0:
            continue
        else:
            what += f" [arg{j}: sizes {ta.numel()} vs {tb.numel()}]"
            continue
        d = float((a.double() - b.double()).norm() / max(float(a.double().norm()), 1e-30))
        if d > worst:
            worst, wj = d, j
    flag = " <<<<" if worst > 1e-6 else ""
    print(f"{i:4d} {fa:32s} worst rel diff {worst:.3e}{what}{flag}")
