"""Host orchestration (kernel chain, strides, gradient accumulation) against the oracle, on CPU.

The C-ABI dispatcher is routed to oracle/abi_reference.py by the `cpu_backend` fixture, so only
the Python host logic of gnnepcsaft_b200 is under test here; the CUDA kernels themselves are
checked against the same per-entry-point reference in tests/test_kernels_gpu.py.
"""
import pytest
import torch

from gnnepcsaft_b200.data.graph import collate
from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
from gnnepcsaft_b200.train import models as M
from oracle import pna_oracle as O
from tests.helpers import CONFIGS, cfg, flat_grads, make_pair, tiny_graphs, zero_grad_bias_mask


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("train", [False, True])
def test_forward_backward_vs_oracle(cpu_backend, name, train):
    c = cfg(**CONFIGS[name])
    if c.hidden_dim >= 64:
        c.propagation_depth = min(c.propagation_depth, 3)
    packed = make_packed_graphs(6, seed=3, num_para=c.num_para)
    batch = batch_from_packed(packed)
    o32, o64, m = make_pair(c, packed.in_degree_histogram())
    for mod in (o32, o64, m):
        mod.train(train)
    tgt = batch.para.view(-1, c.num_para)
    y64 = o64(batch)
    O.mape(y64, tgt.double()).backward()
    y32 = o32(batch)
    ym = m(batch)
    M.mape(ym, tgt).backward()
    scale = float(y64.abs().max())
    err_m = float((ym.double() - y64).abs().max()) / scale
    err_o = float((y32.double() - y64).abs().max()) / scale
    assert err_m <= max(2 * err_o, 1e-5), (err_m, err_o)
    g64, gm = flat_grads(o64), flat_grads(m)
    mask = zero_grad_bias_mask(m)
    rel = float((gm - g64)[mask].norm() / g64[mask].norm())
    assert rel < 2e-3, rel
    if train:   # running statistics / num_batches_tracked must follow torch BatchNorm1d
        for (k, bo), (_, bm) in zip(o32.named_buffers(), m.named_buffers()):
            torch.testing.assert_close(bm.float(), bo.float(), rtol=1e-4, atol=1e-5, msg=k)


def test_tiny_graphs_and_batch_none(cpu_backend):
    c = cfg(propagation_depth=2)
    o32, o64, m = make_pair(c, torch.tensor([1, 6, 8, 2, 1]))
    o64.eval(), m.eval()
    graphs = tiny_graphs()
    with torch.no_grad():
        torch.testing.assert_close(m(collate(graphs)).double(), o64(collate(graphs)), rtol=1e-5, atol=1e-5)
        for g in graphs:
            y = m(g)
            assert y.shape == (1, 5)
            torch.testing.assert_close(y.double(), o64(g), rtol=1e-5, atol=1e-5)


def test_state_dict_roundtrip_and_lightning_prefix(cpu_backend):
    c = cfg(propagation_depth=2)
    o32, _, m = make_pair(c, torch.tensor([0, 3, 2, 1]), dtype64=False)
    assert list(m.state_dict().keys()) == list(o32.state_dict().keys())
    for k, v in o32.state_dict().items():
        assert torch.equal(m.state_dict()[k], v), k
    from gnnepcsaft_b200.train.utils import create_model
    c.model = "PNAL"
    wrapped = create_model(c, torch.tensor([0, 3, 2, 1]))
    sd = {"model." + k: v for k, v in o32.state_dict().items()}
    wrapped.load_state_dict(sd, strict=True)
    o32.load_state_dict(m.state_dict(), strict=True)      # and back


def test_arena_survives_to_and_load(cpu_backend):
    c = cfg(propagation_depth=1)
    _, _, m = make_pair(c, torch.tensor([0, 3, 2, 1]), dtype64=False)
    a = m.arena()
    assert m._arena_valid()
    m.float()                       # nn.Module._apply may re-create parameter storage
    b = m.arena()
    assert m._arena_valid()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m.load_state_dict(sd)
    assert m._arena_valid() and b.data_ptr() == m.arena().data_ptr()
    with pytest.raises(RuntimeError):
        m.double().arena()


def test_cpu_tensors_are_rejected():
    """No CPU path: without the emulator fixture a CPU batch must fail loudly."""
    c = cfg(propagation_depth=1)
    _, _, m = make_pair(c, torch.tensor([0, 3, 2, 1]), dtype64=False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(batch_from_packed(make_packed_graphs(2, seed=0)))


def test_engine_steps_match_reference_training_loop(cpu_backend):
    """TrainEngine.step == reference training_step + AdamW(amsgrad) + CosineAnnealingWarmRestarts per step."""
    from gnnepcsaft_b200.train.engine import TrainEngine
    c = cfg(propagation_depth=2, warmup_steps=3)
    packed = make_packed_graphs(12, seed=4)
    o32, _, m = make_pair(c, packed.in_degree_histogram(), dtype64=False)
    o32.train()
    opt = torch.optim.AdamW(o32.parameters(), lr=c.learning_rate, weight_decay=c.weight_decay, amsgrad=True, eps=1e-5)
    sched = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, c.warmup_steps)
    engine = TrainEngine.from_config(m, c, distributed=False)
    p_start = torch.cat([p.data.reshape(-1).clone() for p in o32.parameters()])
    for it in range(5):
        batch = batch_from_packed(packed, torch.arange(it % 2 * 6, it % 2 * 6 + 6))
        opt.zero_grad()
        loss_ref = O.mape(o32(batch), batch.para.view(-1, c.num_para))
        loss_ref.backward()
        opt.step()
        sched.step()
        loss = engine.step(batch)
        assert abs(float(loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref)) + 1e-6, it
    # Adam normalises each element by its own gradient scale, so elements whose gradient is rounding
    # noise move by up to ~lr in either implementation: compare the update as a whole, not elementwise
    p_ref = torch.cat([p.data.reshape(-1) for p in o32.parameters()])
    p_eng = torch.cat([p.data.reshape(-1) for p in m.parameters()])
    rel = float((p_eng - p_ref).norm() / (p_ref - p_start).norm())
    assert rel < 5e-2, rel
    out = engine.predict(batch_from_packed(packed))
    o32.eval()
    with torch.no_grad():
        torch.testing.assert_close(out, o32(batch_from_packed(packed)), rtol=1e-3, atol=1e-4)
    sd = engine.optimizer_state_dict()
    assert len(sd["state"]) == len(list(m.parameters())) and sd["param_groups"][0]["amsgrad"]


def test_inference_plan_follows_weight_changes(cpu_backend):
    """The eval fast path caches weight-derived tensors (bond tables, folded weights, packed images) in the model's
    inference plan: every way the weights can change must invalidate it."""
    from gnnepcsaft_b200.train.engine import TrainEngine
    from gnnepcsaft_b200.train.utils import create_model
    c = cfg(propagation_depth=2)
    packed = make_packed_graphs(10, seed=4)
    deg = packed.in_degree_histogram()
    batch = batch_from_packed(packed)
    _, _, m = make_pair(c, deg, dtype64=False)
    engine = TrainEngine.from_config(m, c, distributed=False)

    def fresh_prediction():
        ref = create_model(c, deg)
        ref.load_state_dict(m.state_dict())
        ref.eval()
        with torch.no_grad():
            return ref(batch)

    y0 = engine.predict(batch)
    assert m._plan and all(pl.token is not None for pl in m._plan.values())
    torch.testing.assert_close(engine.predict(batch), y0, rtol=0, atol=0)          # served from the plan
    engine.step(batch)                                                              # raw-pointer update (AdamW, BN)
    y1 = engine.predict(batch)
    assert not torch.equal(y1, y0)
    torch.testing.assert_close(y1, fresh_prediction(), rtol=1e-6, atol=1e-7)
    with torch.no_grad():                                                           # in-place edit by the user
        m.convs[0].lin.weight.mul_(1.5)
    torch.testing.assert_close(engine.predict(batch), fresh_prediction(), rtol=1e-6, atol=1e-7)
    sd = {k: v.clone() for k, v in m.state_dict().items()}                          # load_state_dict
    sd["mlp.0.weight"] = sd["mlp.0.weight"] * 0.5
    m.load_state_dict(sd)
    torch.testing.assert_close(engine.predict(batch), fresh_prediction(), rtol=1e-6, atol=1e-7)
    m.train()                                                                       # autograd training forward: BN stats move
    M.mape(m(batch), batch.para.view(-1, c.num_para)).backward()
    m.eval()
    with torch.no_grad():
        torch.testing.assert_close(m(batch), fresh_prediction(), rtol=1e-6, atol=1e-7)
