"""Ensemble batched inference (host logic, C ABI emulated on CPU): [M, K, P] in one pass equals the reference's
per-molecule x per-model loop (evaluate_ensemble.py:128-208) run with the float64 oracle."""
import torch

from gnnepcsaft_b200.data.compact import CompactGraphStore
from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
from gnnepcsaft_b200.train.ensemble import EnsemblePredictor
from tests.helpers import cfg, make_pair


def test_ensemble_matches_per_molecule_per_model_loop(cpu_backend):
    c = cfg(propagation_depth=2)
    packed = make_packed_graphs(9, seed=8, num_para=c.num_para)
    deg = packed.in_degree_histogram()
    pairs = [make_pair(c, deg, seed=s) for s in (0, 1, 2)]
    ens = EnsemblePredictor([m for _, _, m in pairs])
    assert len(ens) == 3
    batch = batch_from_packed(packed)
    y = ens.predict(batch, dtype=torch.float64)
    assert y.shape == (9, 3, c.num_para) and y.dtype == torch.float64
    # the reference: one molecule at a time, one model at a time, float64 models, concat + mean(0)
    for i, g in enumerate(packed.to_list()):
        preds = []
        for _, o64, _ in pairs:
            o64.eval()
            with torch.no_grad():
                preds.append(o64(g))
        ref = torch.concat(preds, dim=0)
        scale = float(ref.abs().max())
        assert float((y[i] - ref).abs().max()) / scale <= 1e-5, i
        assert float((ens.predict_mean(batch, torch.float64)[i] - ref.mean(0)).abs().max()) / scale <= 1e-5
    # store-resident paths give the same numbers
    store = CompactGraphStore(packed, device="cpu")
    ids = torch.tensor([4, 0, 7])
    torch.testing.assert_close(ens.predict_from_store(store, ids), ens.predict(batch_from_packed(packed, ids)),
                               rtol=1e-5, atol=1e-6)
    mean, allp = ens.screen(store, 1, 9, max_nodes=90)
    torch.testing.assert_close(allp, ens.predict(batch)[1:9], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(mean, allp.mean(1))


def test_ensemble_from_state_dicts_accepts_lightning_prefix(cpu_backend):
    c = cfg(propagation_depth=1)
    packed = make_packed_graphs(4, seed=1, num_para=c.num_para)
    deg = packed.in_degree_histogram()
    o32, _, m = make_pair(c, deg, dtype64=False)
    sds = [o32.state_dict(), {"model." + k: v for k, v in o32.state_dict().items()}]
    ens = EnsemblePredictor.from_state_dicts(c, deg, sds, device="cpu")
    y = ens.predict(batch_from_packed(packed))
    torch.testing.assert_close(y[:, 0], y[:, 1])
    m.eval()
    with torch.no_grad():
        torch.testing.assert_close(y[:, 0], m(batch_from_packed(packed)))
