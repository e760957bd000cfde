"""The oracle against its independent float64 loop restatement and the committed golden vectors."""
import os

import numpy as np
import pytest
import torch

from gnnepcsaft_b200.data.graph import collate
from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
from oracle import naive_loops as NL
from oracle import pna_oracle as O
from tests.helpers import CONFIGS, cfg, randomise_bn, tiny_graphs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _naive(model, c, data):
    return NL.forward_eval(model.state_dict(), data.x, data.edge_index, data.edge_attr, data.batch,
                           hidden=c.hidden_dim, depth=c.propagation_depth, pre_layers=c.pre_layers,
                           post_layers=c.post_layers, num_mlp_layers=c.num_mlp_layers, skip=c.skip_connections,
                           self_loops=c.add_self_loops)


@pytest.mark.parametrize("name", ["default", "no_loops_no_skip", "pre2_post2", "pre3_post3"])
def test_oracle_matches_naive_loops(name):
    c = cfg(**CONFIGS[name])
    if c.hidden_dim > 16:
        c.propagation_depth = min(c.propagation_depth, 2)
    graphs = tiny_graphs() + make_packed_graphs(2, seed=5).to_list()
    deg = torch.tensor([1, 8, 6, 3, 2])
    torch.manual_seed(0)
    model = O.create_model(c, deg).double()
    randomise_bn(model)
    model.eval()
    batch = collate(graphs)
    with torch.no_grad():
        y = model(batch).numpy()
    yn = _naive(model, c, batch)
    np.testing.assert_allclose(y, yn, rtol=1e-10, atol=1e-10)


def test_oracle_single_molecule_batch_none():
    c = cfg(propagation_depth=2)
    torch.manual_seed(0)
    model = O.create_model(c, torch.tensor([0, 3, 2, 1])).double().eval()
    for g in tiny_graphs():
        with torch.no_grad():
            y = model(g)
        assert y.shape == (1, 5)
        np.testing.assert_allclose(y.numpy(), _naive(model, c, g), rtol=1e-10, atol=1e-10)


def test_oracle_edge_semantics():
    """deg-0 node without self-loops -> all aggregates 0; exact ties split the min/max gradient evenly."""
    agg = O.DegreeScalerAggregation(torch.tensor([0, 1, 2]))
    m = torch.tensor([[[1.0, 2.0]], [[1.0, 2.0]], [[3.0, 0.0]]], requires_grad=True)   # [E=3, T=1, F=2]
    out = agg(m, torch.tensor([0, 0, 2]), 3)                                         # node 1 has no in-edge
    assert torch.all(out[1] == 0)
    # node 0: identical messages -> var = 0 -> std masked to 0
    assert torch.all(out[0, 0, 6:8] == 0)
    out[0, 0, 2].backward()                      # min of channel 0 at node 0, two tied edges
    assert m.grad[0, 0, 0] == 0.5 and m.grad[1, 0, 0] == 0.5


def test_golden_vectors():
    """Outputs committed by oracle/make_golden.py are reproduced by the oracle in this environment."""
    path = os.path.join(GOLDEN, "pna_golden.pt")
    assert os.path.exists(path), "run python oracle/make_golden.py"
    gold = torch.load(path)
    from oracle.make_golden import CASES, build_case
    for name in CASES:
        model, batch, _ = build_case(name)
        model.eval()
        with torch.no_grad():
            y = model(batch)
        torch.testing.assert_close(y, gold[name]["eval_out_fp32"], rtol=1e-5, atol=1e-6)
