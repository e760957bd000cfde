"""Shared builders for the tests."""
import types

import torch

from gnnepcsaft_b200.data.graph import Data, collate
from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
from gnnepcsaft_b200.train.utils import create_model, default_config
from oracle import pna_oracle as O


def cfg(**kw):
    c = default_config()
    c.model = "PNA"
    for k, v in kw.items():
        setattr(c, k, v)
    return c


CONFIGS = {
    "default": dict(),
    "no_loops_no_skip": dict(add_self_loops=False, skip_connections=False, propagation_depth=2),
    "pre2_post2": dict(pre_layers=2, post_layers=2, num_mlp_layers=0, propagation_depth=2, hidden_dim=16,
                       add_self_loops=False, skip_connections=False),
    "pre3_post3": dict(pre_layers=3, post_layers=3, num_mlp_layers=2, propagation_depth=2, hidden_dim=8,
                       num_para=3),
    "wide_small": dict(hidden_dim=128, propagation_depth=2, post_layers=2, add_self_loops=False),
}


def tiny_graphs():
    """1-atom molecule (E=0), symmetric 3-chain (exact min/max ties), 6-ring."""
    g0 = Data(x=torch.tensor([[6, 0, 4, 5, 4, 0, 2, 0, 0]]), edge_index=torch.zeros(2, 0, dtype=torch.int64),
              edge_attr=torch.zeros(0, 3, dtype=torch.int64), para=torch.tensor([1.0, 3.0, 200.0, 0.5, 1000.0]))
    x1 = torch.tensor([[7, 0, 1, 5, 2, 0, 2, 0, 0], [5, 0, 2, 5, 2, 0, 2, 0, 0], [7, 0, 1, 5, 2, 0, 2, 0, 0]])
    e1 = torch.tensor([[0, 1, 1, 2], [1, 0, 2, 1]])
    g1 = Data(x=x1, edge_index=e1, edge_attr=torch.zeros(4, 3, dtype=torch.int64),
              para=torch.tensor([2.0, 3.5, 250.0, 0.7, 1500.0]))
    x2 = torch.tensor([[5, 0, 3, 5, 1, 0, 1, 1, 1]] * 6)
    ring = [(i, (i + 1) % 6) for i in range(6)]
    e2 = torch.tensor([[a for a, b in ring for _ in (0, 1)], [b for a, b in ring for _ in (0, 1)]])
    e2 = torch.stack([torch.tensor([v for a, b in ring for v in (a, b)]),
                      torch.tensor([v for a, b in ring for v in (b, a)])])
    g2 = Data(x=x2, edge_index=e2, edge_attr=torch.tensor([[3, 0, 1]] * 12),
              para=torch.tensor([2.5, 3.6, 280.0, 0.9, 1800.0]))
    return [g0, g1, g2]


def randomise_bn(model, seed=1):
    g = torch.Generator().manual_seed(seed)
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.5)
            mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 1.5 + 0.5)
            mod.weight.data.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
            mod.bias.data.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)


def make_pair(config, deg, seed=0, dtype64=True):
    """(oracle fp32, oracle fp64, product model) with identical weights."""
    torch.manual_seed(seed)
    o32 = O.create_model(config, deg)
    randomise_bn(o32)
    o64 = None
    if dtype64:
        o64 = O.create_model(config, deg).double()
        o64.load_state_dict(o32.state_dict())
    m = create_model(config, deg)
    m.load_state_dict(o32.state_dict(), strict=True)
    return o32, o64, m


def flat_grads(model):
    return torch.cat([p.grad.reshape(-1).double().cpu() for p in model.parameters()])


def zero_grad_bias_mask(model):
    """True for live parameters.  Biases whose output reaches a training-mode BatchNorm through
    linear maps only have a mathematically zero gradient (hard part 2)."""
    pp, mp = model.pna_params, model.mlp_params
    dead = set()
    for l in range(pp.propagation_depth):
        dead.add(f"convs.{l}.lin.bias")
        for t in range(2):
            dead.add(f"convs.{l}.post_nns.{t}.{2 * (pp.post_layers - 1)}.bias")
    for i in range(mp.num_mlp_layers):
        dead.add(f"mlp.{4 * i}.bias")
    k = 4 * mp.num_mlp_layers
    dead.update({f"mlp.{k}.0.bias", f"mlp.{k}.4.bias"})
    mask = []
    for name, p in model.named_parameters():
        mask.append(torch.full((p.numel(),), name not in dead))
    return torch.cat(mask)


# bf16 node-update GEMMs (matmul_precision = "bf16", the reference's own torch "medium" matmul precision,
# train.py:319): (forward max-normalised error vs the fp64 oracle, gradient global rel-L2) bars = 2 x the values
# measured on B200 (listed in DESIGN.md section 2); the projection GEMM stays TF32x3 in this mode.
# Measured on B200 (round 2): default 7.5e-3 / 0.224, pre2_post2 1.1e-2 / 0.174, wide_small 7.4e-3 / 0.216,
# wide (H 256, L 6) 2.4e-2 (train-mode forward) / 0.300.  The gradient figures are what bf16 operand rounding does to
# THIS architecture (min / max / std selections flip): the fp64 oracle with the same operands rounded to bf16 measures
# 0.218 / 0.131 / 0.189 / 0.242 - the tests additionally require the CUDA path to stay within 2x that floor.
BF16_BARS = {
    "default": (1.5e-2, 4.5e-1),
    "pre2_post2": (2.2e-2, 3.5e-1),
    "wide_small": (1.5e-2, 4.4e-1),
    "wide": (5e-2, 6e-1),
}


class bf16_operand_rounding:
    """Context manager: the Linear layers that `matmul_precision = "bf16"` runs with bf16 operands (post_nns, lin,
    the extra pre_nns layers and the MLP head - not the first pre_nns layer / edge_encoder, whose projections stay
    TF32x3) round their input and weight to bf16 and accumulate in the module's dtype.  Applied to the fp64 oracle
    this isolates the effect of bf16 operand rounding on the reference arithmetic: the floor any bf16-operand
    implementation sits on, against which the CUDA bf16 path is held (factor 2)."""

    def __init__(self, model):
        self.model, self.saved = model, []

    @staticmethod
    def _rounded(lin):
        def fwd(x):
            xr = x.to(torch.bfloat16).to(x.dtype)
            wr = lin.weight.to(torch.bfloat16).to(lin.weight.dtype)
            return torch.nn.functional.linear(xr, wr, lin.bias)
        return fwd

    def __enter__(self):
        for name, mod in self.model.named_modules():
            if not isinstance(mod, torch.nn.Linear):
                continue
            first_pre = ".pre_nns." in name and name.endswith(".0")
            if "edge_encoder" in name or first_pre:
                continue
            self.saved.append(mod)
            mod.forward = self._rounded(mod)
        return self.model

    def __exit__(self, *exc):
        for mod in self.saved:
            del mod.forward
        return False
