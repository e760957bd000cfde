"""Per-parameter gradient error of the CUDA model vs the fp64 oracle (diagnostic; run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-epc-saft_b200")):
    sys.path.insert(0, p)
import torch
from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
from gnnepcsaft_b200.train import models as M
from gnnepcsaft_b200.train.utils import wide_config
from oracle import pna_oracle as O
from tests.helpers import make_pair, zero_grad_bias_mask, cfg

which = sys.argv[1] if len(sys.argv) > 1 else "wide"
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["simt", "tf32x3"]
if which.startswith("wide"):
    c = wide_config(); c.model = "PNA"
    if which == "wide_loops":
        c.add_self_loops = True
    if which == "wide_post1":
        c.post_layers = 1
    if which == "wide_d2":
        c.propagation_depth = 2
    packed = make_packed_graphs(160, seed=4, skewed=True)
else:
    c = cfg(); packed = make_packed_graphs(512, seed=4)
batch = batch_from_packed(packed)
o32, o64, m = make_pair(c, packed.in_degree_histogram())
for mod in (o32, o64, m):
    mod.train()
tgt = batch.para.view(-1, c.num_para)
O.mape(o64(batch), tgt.double()).backward()
O.mape(o32(batch), tgt).backward()
m.to("cuda")
names = [n for n, _ in m.named_parameters()]
g64 = {n: p.grad.double() for n, p in o64.named_parameters()}
g32 = {n: p.grad.double() for n, p in o32.named_parameters()}
for mode in modes:
    m.matmul_precision = mode
    m.zero_grad()
    M.mape(m(batch.to("cuda")), tgt.cuda()).backward()
    g = {n: p.grad.double().cpu() for n, p in m.named_parameters()}
    tot = sum(float((g[n] - g64[n]).norm() ** 2) for n in names) ** 0.5 / sum(float(g64[n].norm() ** 2) for n in names) ** 0.5
    tot32 = sum(float((g32[n] - g64[n]).norm() ** 2) for n in names) ** 0.5 / sum(float(g64[n].norm() ** 2) for n in names) ** 0.5
    print(f"== {which}/{mode}: global rel-L2 {tot:.3e} (oracle32 {tot32:.3e})")
    rows = []
    for n in names:
        den = float(g64[n].norm())
        if den == 0:
            continue
        rows.append((float((g[n] - g64[n]).norm()), float((g32[n] - g64[n]).norm()), den, n))
    rows.sort(reverse=True)
    if os.environ.get("ALL"):
        for n in names:
            den = float(g64[n].norm())
            if den > 0:
                print(f"   {n:45s} rel {float((g[n] - g64[n]).norm()) / den:.2e} (oracle32 {float((g32[n] - g64[n]).norm()) / den:.2e}) |g| {den:.2e}")
    else:
        for e, e32, den, n in rows[:14]:
            print(f"   {n:45s} |err| {e:.3e} (oracle32 {e32:.3e})  |g| {den:.3e}  rel {e / den:.2e}")
