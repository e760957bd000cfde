"""Per-parameter gradient differences: padded compact batch vs PyG-layout batch vs reordered batch (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-epc-saft_b200")):
    sys.path.insert(0, p)
import torch
from gnnepcsaft_b200.data.compact import CompactGraphStore
from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
from gnnepcsaft_b200.train.engine import TrainEngine
from tests.helpers import cfg, make_pair

DEV = "cuda"
c = cfg()
packed = make_packed_graphs(300, seed=2, num_para=c.num_para)
deg = packed.in_degree_histogram()
ids = torch.randperm(300, generator=torch.Generator().manual_seed(0))[:256]
store = CompactGraphStore(packed, device=DEV, self_loops=c.add_self_loops)
grads = {}
for name in ("plain", "reorder", "padded", "exact"):
    _, _, m = make_pair(c, deg, dtype64=False)
    m.to(DEV)
    e = TrainEngine.from_config(m, c, distributed=False)
    e.use_cuda_graphs = False
    if name == "plain":
        e.step(batch_from_packed(packed, ids).to(DEV))
    elif name == "reorder":
        e.step(batch_from_packed(packed, ids.flip(0)).to(DEV))
    elif name == "padded":
        e.step_from_store(store, ids)
    else:
        e.step(store.batch(ids, pad=False))
    torch.cuda.synchronize()
    grads[name] = e.grad[:e.numel].double().cpu()
    model = m
g1 = grads["plain"]
for k in ("reorder", "padded", "exact"):
    print(k, float((grads[k] - g1).norm() / g1.norm()))
rows = []
for (name, p), off in zip(model.named_parameters(), model._arena_offsets):
    sl = slice(off, off + p.numel())
    a = g1[sl]
    rows.append((float((grads["padded"][sl] - a).norm()), float((grads["reorder"][sl] - a).norm()),
                 float((grads["exact"][sl] - a).norm()), float(a.norm()), name))
rows.sort(reverse=True)
for r in rows[:14]:
    print("pad %.3e  reorder %.3e  exact %.3e  |g| %.3e  %s" % r)

# leak check: the gradient must not depend on what the padding nodes contain, nor on how much padding there is
from gnnepcsaft_b200.data.compact import bucket
n, s = store.sizes(ids)
res = {}
for name, pad, scramble in (("pad_a", bucket(n, s), False), ("pad_a_scrambled", bucket(n, s), True),
                            ("pad_big", (n + 700, s + 3000), False)):
    _, _, m = make_pair(c, deg, dtype64=False)
    m.to(DEV)
    e = TrainEngine.from_config(m, c, distributed=False)
    e.use_cuda_graphs = False
    cb = store.batch(ids, pad=pad)
    if scramble:
        cb.x[n:] = torch.randint(0, 2, cb.x[n:].shape, device=DEV, dtype=torch.uint8)
    e.step(cb)
    torch.cuda.synchronize()
    res[name] = e.grad[:e.numel].double().cpu()
for k in ("pad_a_scrambled", "pad_big"):
    print(k, "vs pad_a", float((res[k] - res["pad_a"]).norm() / res["pad_a"].norm()))

gp, gx = res["pad_a"], grads["exact"]
print("projection coefficient pad on exact - 1:", float((gp @ gx) / (gx @ gx) - 1))
rows = []
for (name, p), off in zip(model.named_parameters(), model._arena_offsets):
    sl = slice(off, off + p.numel())
    a, b = gx[sl], gp[sl]
    if float(a.norm()) > 0:
        rows.append((float((b - a).norm() / a.norm()), float((b @ a) / (a @ a) - 1), float(a.norm()), name))
rows.sort(reverse=True)
for r in rows[:40]:
    print("rel %.3e  coef-1 %.3e  |g| %.3e  %s" % r)

# which cluster is closer to the truth?  fp64 oracle gradient of the same step
from oracle import pna_oracle as O
from tests.helpers import flat_grads, zero_grad_bias_mask
o32, o64, m = make_pair(c, deg, dtype64=True)
o64.train()
hb = batch_from_packed(packed, ids)
O.mape(o64(hb), hb.para.view(-1, c.num_para).double()).backward()
# arena layout pads every tensor to a multiple of 4 floats: gather per parameter
def per_param(gflat):
    return torch.cat([gflat[off:off + p.numel()] for p, off in zip(model.parameters(), model._arena_offsets)])
g64 = flat_grads(o64)
mask = zero_grad_bias_mask(m)
for k, gv in (("plain", grads["plain"]), ("exact", grads["exact"]), ("reorder", grads["reorder"]), ("pad_a", res["pad_a"]), ("pad_big", res["pad_big"])):
    gg = per_param(gv)
    print(f"{k:8s} vs fp64 oracle: rel-L2 {float((gg - g64)[mask].norm() / g64[mask].norm()):.3e}")
print("n =", n, "n+700 =", n + 700, "bucket", bucket(n, s))
