"""ORACLE (test infrastructure): writes tests/golden/pna_golden.pt.

The reference cannot be imported here (torch_geometric / ogb absent), so
the golden vectors are outputs of oracle/pna_oracle.py (PARITY UNPINNED) on seeded synthetic
batches: eval-mode outputs in fp32 and fp64, train-mode outputs, the MAPE loss and the flattened
parameter gradient (fp64).  Run from the repo root:  python oracle/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-epc-saft_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from gnnepcsaft_b200.data.graph import collate  # noqa: E402
from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs  # noqa: E402
from oracle import pna_oracle as O  # noqa: E402
from tests.helpers import CONFIGS, cfg, randomise_bn, tiny_graphs  # noqa: E402

CASES = {
    "tiny_default": ("default", "tiny", 0),
    "default_256": ("default", 256, 1),
    "no_loops_no_skip_64": ("no_loops_no_skip", 64, 2),
    "pre2_post2_64": ("pre2_post2", 64, 3),
    "wide_small_skewed_48": ("wide_small", -48, 4),
}


def build_case(name, dtype=torch.float32):
    cname, size, seed = CASES[name]
    c = cfg(**CONFIGS[cname])
    if size == "tiny":
        graphs = tiny_graphs()
        batch = collate(graphs)
        deg = torch.tensor([1, 6, 8, 2, 1])
    else:
        packed = make_packed_graphs(abs(size), seed=seed, skewed=size < 0, num_para=c.num_para)
        batch = batch_from_packed(packed)
        deg = packed.in_degree_histogram()
    torch.manual_seed(seed)
    model = O.create_model(c, deg)
    randomise_bn(model, seed=seed + 1)
    if dtype == torch.float64:
        model = model.double()
    return model, batch, c


def grad_projections(flat, k=8):
    g = torch.Generator().manual_seed(1234)
    probes = torch.randn(k, flat.numel(), generator=g, dtype=torch.float64)
    return probes @ flat.double()


def main():
    out = {}
    for name in CASES:
        rec = {}
        m32, batch, c = build_case(name)
        m64 = O.create_model(c, m32.pna_params.deg).double()
        m64.load_state_dict(m32.state_dict())
        target = batch.para.view(-1, c.num_para)
        for tag, m, tgt in (("fp32", m32, target), ("fp64", m64, target.double())):
            m.eval()
            with torch.no_grad():
                rec[f"eval_out_{tag}"] = m(batch).clone()
        m64.train()
        y = m64(batch)
        loss = O.mape(y, target.double())
        loss.backward()
        rec["train_out_fp64"] = y.detach().clone()
        rec["train_loss_fp64"] = loss.detach().clone()
        # the full gradient is large; pin it by per-parameter norms and seeded random projections
        flat = torch.cat([p.grad.reshape(-1) for p in m64.parameters()])
        rec["train_grad_norms_fp64"] = torch.stack([p.grad.norm() for p in m64.parameters()])
        rec["train_grad_proj_fp64"] = grad_projections(flat)
        out[name] = rec
    path = os.path.join(ROOT, "tests", "golden", "pna_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
