"""Model-level parity at the REAL BASELINE.json shapes ("What's weak" #2).

config 2: default model, 4096-graph fp32 TRAIN step (forward + MAPE + backward) against the fp64 oracle, and
          TrainEngine.step against the oracle loss;
config 5: the wide model of reference configs/configs_parallel.py:43-51 (hidden 256, 6 PNAConv, post_layers 2,
          no self-loops) on skewed 5-150-atom graphs, eval + train, for the fp32-parity tensor-core mode (tf32x3)
          and for bf16 node-update GEMMs, the latter held to 2x the values measured on B200 (tests/helpers.py
          BF16_BARS, also listed in DESIGN.md).
Truth is the fp64 oracle; bars follow SURVEY.md section 8c (train-mode forward floor 1.4e-4 from std-threshold flips,
gradient global rel-L2 2e-3, mathematically-zero biases excluded)."""
import pytest
import torch

from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
from gnnepcsaft_b200.train import models as M
from gnnepcsaft_b200.train.utils import wide_config
from oracle import pna_oracle as O
from tests.helpers import BF16_BARS, bf16_operand_rounding, cfg, flat_grads, make_pair, zero_grad_bias_mask

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _norm_err(a, b):
    return float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max())


def _train_pass(model, batch, num_para, dev=None):
    model.zero_grad()
    b = batch.to(dev) if dev else batch
    tgt = b.para.view(-1, num_para)
    if next(model.parameters()).dtype == torch.float64:
        tgt = tgt.double()
    y = model(b)
    loss = (M.mape if dev else O.mape)(y, tgt)
    loss.backward()
    return y.detach(), float(loss)


def test_config2_train_step_4096_graphs_vs_fp64_oracle():
    c = cfg()
    packed = make_packed_graphs(4096, seed=1)
    batch = batch_from_packed(packed)
    o32, o64, m = make_pair(c, packed.in_degree_histogram())
    for mod in (o32, o64, m):
        mod.train()
    m.to(DEV)
    y64, l64 = _train_pass(o64, batch, c.num_para)
    y32, l32 = _train_pass(o32, batch, c.num_para)
    y, l = _train_pass(m, batch, c.num_para, DEV)
    e_m, e_o = _norm_err(y, y64), _norm_err(y32, y64)
    mask = zero_grad_bias_mask(m)
    g64, g32, g = flat_grads(o64), flat_grads(o32), flat_grads(m)
    rel = float((g - g64)[mask].norm() / g64[mask].norm())
    rel_o = float((g32 - g64)[mask].norm() / g64[mask].norm())
    print(f"config2 4096-graph train: fwd cuda-vs-fp64 {e_m:.2e} (oracle32 {e_o:.2e}); loss {l:.7f} vs {l64:.7f}; "
          f"grad rel-L2 cuda {rel:.2e} (oracle32 {rel_o:.2e})")
    assert e_m <= max(2 * e_o, 2e-4)
    assert abs(l - l64) <= 1e-5 * abs(l64)
    assert rel <= max(2e-3, 2 * rel_o)
    for (k, bo), (_, bm) in zip(o64.named_buffers(), m.named_buffers()):
        torch.testing.assert_close(bm.double().cpu(), bo.double(), rtol=1e-4, atol=1e-5, msg=k)

    # the trainer-free engine computes the same loss on the same batch (fresh model: same running statistics)
    from gnnepcsaft_b200.train.engine import TrainEngine
    _, _, m2 = make_pair(c, packed.in_degree_histogram(), dtype64=False)
    eng = TrainEngine.from_config(m2.to(DEV), c, distributed=False)
    le = float(eng.step(batch.to(DEV)))
    assert abs(le - l64) <= 1e-5 * abs(l64), (le, l64)


@pytest.mark.parametrize("precision", ["tf32x3", "bf16"])
def test_config5_wide_model_skewed_graphs(precision):
    c = wide_config()
    c.model = "PNA"
    assert (c.hidden_dim, c.propagation_depth, c.post_layers, c.add_self_loops) == (256, 6, 2, False)
    packed = make_packed_graphs(160, seed=4, skewed=True)
    batch = batch_from_packed(packed)
    o32, o64, m = make_pair(c, packed.in_degree_histogram())
    m.matmul_precision = precision
    m.to(DEV)
    for mod in (o32, o64, m):
        mod.eval()
    with torch.no_grad():
        y64, y32, y = o64(batch), o32(batch), m(batch.to(DEV))
        with bf16_operand_rounding(o64):        # (before any train-mode pass touches the running statistics)
            f_e = _norm_err(o64(batch), y64)
    e_m, e_o = _norm_err(y, y64), _norm_err(y32, y64)
    for mod in (o32, o64, m):
        mod.train()
    yt64, l64 = _train_pass(o64, batch, c.num_para)
    yt32, _ = _train_pass(o32, batch, c.num_para)
    yt, l = _train_pass(m, batch, c.num_para, DEV)
    t_m, t_o = _norm_err(yt, yt64), _norm_err(yt32, yt64)
    mask = zero_grad_bias_mask(m)
    g64, g32, g = flat_grads(o64), flat_grads(o32), flat_grads(m)
    rel = float((g - g64)[mask].norm() / g64[mask].norm())
    rel_o = float((g32 - g64)[mask].norm() / g64[mask].norm())
    print(f"config5 wide/{precision}: eval fwd {e_m:.2e} (oracle32 {e_o:.2e}); train fwd {t_m:.2e} (oracle32 {t_o:.2e}); "
          f"loss {l:.6f} vs {l64:.6f}; grad rel-L2 {rel:.2e} (oracle32 {rel_o:.2e})")
    if precision == "bf16":
        # floor: the fp64 oracle with the same Linear operands rounded to bf16
        with bf16_operand_rounding(o64):
            ytb, _ = _train_pass(o64, batch, c.num_para)
            f_t = _norm_err(ytb, yt64)
            f_g = float((flat_grads(o64) - g64)[mask].norm() / g64[mask].norm())
        print(f"config5 wide/bf16 floor (fp64 oracle, bf16-rounded operands): eval fwd {f_e:.2e}, train fwd {f_t:.2e}, "
              f"grad rel-L2 {f_g:.2e}")
        fwd_bar, grad_bar = BF16_BARS["wide"]
        assert e_m <= fwd_bar and t_m <= fwd_bar and rel <= grad_bar
        assert e_m <= 2 * f_e and t_m <= 2 * f_t and rel <= 2 * f_g, "bf16 path worse than 2x the operand-rounding floor"
    else:
        assert e_m <= max(1e-5, 2 * e_o + 1e-6)
        assert t_m <= max(2 * t_o, 2e-4)
        # the wide model's fp32-vs-fp64 gradient floor is set by min / max / std selection flips and is chaotic: on four
        # variants of this config (with / without self-loops, post_layers 1 / 2, depth 2 / 6) the reference arithmetic in
        # fp32 (oracle32) measured 0.9e-3 ... 4.2e-3 against fp64, the CUDA path 1.0e-3 ... 4.1e-3 (tools/grad_diag.py)
        assert rel <= max(5e-3, 2 * rel_o)
