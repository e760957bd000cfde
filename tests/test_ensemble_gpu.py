"""Ensemble batched inference on the GPU: one pass over a batch for K models == the reference's per-molecule x
per-model loop with float64 models (evaluate_ensemble.py:128-208), within the single-model parity bar (1e-5)."""
import pytest
import torch

from gnnepcsaft_b200.data.compact import CompactGraphStore
from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
from gnnepcsaft_b200.train.ensemble import EnsemblePredictor
from tests.helpers import cfg, make_pair

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_ensemble_gpu_matches_float64_reference_loop():
    c = cfg()
    packed = make_packed_graphs(600, seed=21, num_para=c.num_para)
    deg = packed.in_degree_histogram()
    pairs = [make_pair(c, deg, seed=s) for s in (0, 1, 2, 3)]
    ens = EnsemblePredictor([m.to(DEV) for _, _, m in pairs])
    batch = batch_from_packed(packed)
    y = ens.predict(batch.to(DEV), dtype=torch.float64).cpu()
    assert y.shape == (600, 4, c.num_para)
    refs = []
    for _, o64, _ in pairs:
        o64.eval()
        with torch.no_grad():
            refs.append(o64(batch))
    ref = torch.stack(refs, 1)
    err = float((y - ref).abs().max() / ref.abs().max())
    print(f"ensemble [600, 4, 5] vs float64 oracle: max-normalised error {err:.2e}")
    assert err <= 1e-5
    # per molecule, exactly as the reference loops (batch=None graphs)
    for i in (0, 17, 599):
        g = packed.get(i)
        one = torch.concat([o64(g) for _, o64, _ in pairs], dim=0).detach()
        assert float((y[i] - one).abs().max() / one.abs().max()) <= 1e-5
    mean = ens.predict_mean(batch.to(DEV)).cpu()
    assert float((mean.double() - ref.mean(1)).abs().max() / ref.abs().max()) <= 1e-5
    # store-resident: shape buckets + graph replay, and the node-balanced stream
    store = CompactGraphStore(packed, device=DEV)
    ids = torch.randperm(600, generator=torch.Generator().manual_seed(0))[:256]
    for _ in range(3):          # eager, capture + replay, replay
        ys = ens.predict_from_store(store, ids).cpu()
        assert float((ys.double() - ref[ids]).abs().max() / ref.abs().max()) <= 1e-5
    assert ens._engine.graph_replays >= 2
    smean, sall = ens.screen(store, 50, 550, max_nodes=4096)
    assert float((sall.cpu().double() - ref[50:550]).abs().max() / ref.abs().max()) <= 1e-5
    assert torch.equal(smean, sall.mean(1))
