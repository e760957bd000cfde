"""PNAPCSAFT on the CUDA kernels against the oracle (SURVEY.md section 8c parity protocol).

Truth is the float64 oracle.  Forward: max|d| / max|y| <= 1e-5 vs the fp32 oracle and
err(cuda, fp64) <= 2 * err(oracle32, fp64) (+1e-6 slack).  Gradients: global relative L2 <= 2e-3
(the fp32-vs-fp64 floor of the reference arithmetic itself is 6-8e-4), zero-gradient biases
excluded.  Train-mode forward may additionally show std-threshold flips (floor 1.4e-4)."""
import os

import pytest
import torch

from gnnepcsaft_b200.data.graph import collate
from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
from gnnepcsaft_b200.train import models as M
from oracle import pna_oracle as O
from tests.helpers import BF16_BARS, CONFIGS, bf16_operand_rounding, cfg, flat_grads, make_pair, tiny_graphs, zero_grad_bias_mask

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "pna_golden.pt")


def _errs(y, y32, y64):
    scale = float(y64.abs().max())
    return (float((y.double().cpu() - y64).abs().max()) / scale, float((y32.double() - y64).abs().max()) / scale,
            float((y.double().cpu() - y32.double()).abs().max()) / scale)


# fp32 parity bars per dense-stage kernel family: both must be within 1e-5 (max-normalised) of the fp32
# oracle and of the fp64 truth; the exact-fp32 SIMT kernels must additionally be as close to the truth as
# the fp32 oracle itself (factor 2), TF32x3 (tf32 hi+lo split on the tensor cores) is allowed 1e-5.
FP32_MODES = ["tf32x3", "simt"]


def _fp32_bar(mode, e_o):
    return 2 * e_o + 1e-6 if mode == "simt" else 1e-5


@pytest.mark.parametrize("mode", FP32_MODES)
@pytest.mark.parametrize("name", list(CONFIGS))
def test_forward_eval_parity(name, mode):
    c = cfg(**CONFIGS[name])
    packed = make_packed_graphs(200, seed=7, skewed=name == "wide_small", num_para=c.num_para)
    batch = batch_from_packed(packed)
    o32, o64, m = make_pair(c, packed.in_degree_histogram())
    for mod in (o32, o64, m):
        mod.eval()
    m.matmul_precision = mode
    m.to(DEV)
    with torch.no_grad():
        y32, y64, y = o32(batch), o64(batch), m(batch.to(DEV))
    e_m, e_o, e_mo = _errs(y, y32, y64)
    print(f"{name}/{mode}: cuda-vs-fp64 {e_m:.2e} oracle32-vs-fp64 {e_o:.2e} cuda-vs-oracle32 {e_mo:.2e}")
    assert y.shape == (200, c.num_para) and y.dtype == torch.float32
    assert e_mo <= 1e-5, f"vs fp32 oracle {e_mo:.2e}"
    assert e_m <= _fp32_bar(mode, e_o), f"cuda vs fp64 {e_m:.2e}, oracle32 vs fp64 {e_o:.2e}"


@pytest.mark.parametrize("mode", FP32_MODES)
def test_forward_eval_config1_1024_graphs(mode):
    """BASELINE config 1: default model, 1024-graph batch, eval forward."""
    c = cfg()
    packed = make_packed_graphs(1024, seed=0)
    batch = batch_from_packed(packed)
    o32, o64, m = make_pair(c, packed.in_degree_histogram())
    for mod in (o32, o64, m):
        mod.eval()
    m.matmul_precision = mode
    m.to(DEV)
    with torch.no_grad():
        y32, y64, y = o32(batch), o64(batch), m(batch.to(DEV))
    e_m, e_o, e_mo = _errs(y, y32, y64)
    rms = float(((y.double().cpu() - y64) ** 2).mean().sqrt() / (y64 ** 2).mean().sqrt())
    print(f"config1/{mode}: cuda-vs-fp64 {e_m:.2e} (oracle32-vs-fp64 {e_o:.2e}), cuda-vs-oracle32 {e_mo:.2e}, rms {rms:.2e}")
    assert e_mo <= 1e-5 and rms <= 1e-5 and e_m <= _fp32_bar(mode, e_o)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_train_forward_backward_parity(name):
    c = cfg(**CONFIGS[name])
    packed = make_packed_graphs(96, seed=5, num_para=c.num_para)
    batch = batch_from_packed(packed)
    o32, o64, m = make_pair(c, packed.in_degree_histogram())
    for mod in (o32, o64, m):
        mod.train()
    m.to(DEV)
    tgt = batch.para.view(-1, c.num_para)
    y64 = o64(batch)
    O.mape(y64, tgt.double()).backward()
    y32 = o32(batch)
    O.mape(y32, tgt).backward()
    y = m(batch.to(DEV))
    loss = M.mape(y, tgt.to(DEV))
    loss.backward()
    e_m, e_o, _ = _errs(y.detach(), y32.detach(), y64.detach())
    assert e_m <= max(2 * e_o, 2e-4), f"train forward: cuda vs fp64 {e_m:.2e}, oracle32 vs fp64 {e_o:.2e}"
    mask = zero_grad_bias_mask(m)
    g64, g32, g = flat_grads(o64), flat_grads(o32), flat_grads(m)
    rel = float((g - g64)[mask].norm() / g64[mask].norm())
    rel_o = float((g32 - g64)[mask].norm() / g64[mask].norm())
    print(f"{name}: grad rel-L2 cuda {rel:.2e} oracle32 {rel_o:.2e}")
    assert rel <= max(2e-3, 2 * rel_o)
    for (k, bo), (_, bm) in zip(o32.named_buffers(), m.named_buffers()):     # BN running stats, counters
        torch.testing.assert_close(bm.float().cpu(), bo.float(), rtol=1e-4, atol=1e-5, msg=k)


def test_golden_vectors():
    gold = torch.load(GOLDEN)
    from oracle.make_golden import CASES, build_case, grad_projections
    for name in CASES:
        o32, batch, c = build_case(name)
        m = M.PNAPCSAFT(c.hidden_dim, o32.pna_params, o32.mlp_params) if False else None
        from gnnepcsaft_b200.train.utils import create_model
        m = create_model(c, o32.pna_params.deg)
        m.load_state_dict(o32.state_dict(), strict=True)
        m.to(DEV).eval()
        with torch.no_grad():
            y = m(batch.to(DEV)).cpu()
        g64 = gold[name]["eval_out_fp64"]
        scale = float(g64.abs().max())
        err = float((y.double() - g64).abs().max()) / scale
        print(f"golden {name}: eval err {err:.2e}")
        assert err <= 1e-5, (name, err)
        m.train()
        tgt = batch.para.view(-1, c.num_para).to(DEV)
        yt = m(batch.to(DEV))
        loss = M.mape(yt, tgt)
        loss.backward()
        assert abs(float(loss) - float(gold[name]["train_loss_fp64"])) <= 2e-4 * abs(float(gold[name]["train_loss_fp64"]))
        flat = flat_grads(m)
        proj, gproj = grad_projections(flat), gold[name]["train_grad_proj_fp64"]
        perr = float((proj - gproj).norm() / gproj.norm())
        print(f"golden {name}: loss {float(loss):.6f} vs {float(gold[name]['train_loss_fp64']):.6f}, grad-projection err {perr:.2e}")
        assert perr <= 5e-3, (name, perr)


def test_tiny_graphs_edge_cases_and_batch_none():
    c = cfg(propagation_depth=3)
    o32, o64, m = make_pair(c, torch.tensor([1, 6, 8, 2, 1]))
    o64.eval(), m.eval()
    m.to(DEV)
    graphs = tiny_graphs()
    with torch.no_grad():
        torch.testing.assert_close(m(collate(graphs).to(DEV)).double().cpu(), o64(collate(graphs)), rtol=1e-5, atol=1e-5)
        for g in graphs:                              # single molecule, data.batch is None
            y = m(g.to(DEV))
            assert y.shape == (1, 5)
            torch.testing.assert_close(y.double().cpu(), o64(g), rtol=1e-5, atol=1e-5)
    # isolated atoms without self-loops: deg-0 segments
    c2 = cfg(propagation_depth=2, add_self_loops=False)
    o32, o64, m = make_pair(c2, torch.tensor([1, 6, 8, 2, 1]))
    o64.eval(), m.eval()
    m.to(DEV)
    with torch.no_grad():
        torch.testing.assert_close(m(collate(graphs).to(DEV)).double().cpu(), o64(collate(graphs)), rtol=1e-5, atol=1e-5)


def test_state_dict_contract_on_device():
    c = cfg(propagation_depth=2)
    o32, _, m = make_pair(c, torch.tensor([0, 3, 2, 1]), dtype64=False)
    m.to(DEV)
    sd = m.state_dict()
    assert list(sd.keys()) == list(o32.state_dict().keys())
    o32.load_state_dict({k: v.cpu() for k, v in sd.items()}, strict=True)
    packed = make_packed_graphs(8, seed=1)
    m.eval()
    with torch.no_grad():
        y1 = m(batch_from_packed(packed).to(DEV))
        m.load_state_dict({k: v.clone() for k, v in sd.items()})
        y2 = m(batch_from_packed(packed).to(DEV))
    assert torch.equal(y1, y2)


def test_no_cpu_path_and_loud_failure():
    c = cfg(propagation_depth=1)
    _, _, m = make_pair(c, torch.tensor([0, 3, 2, 1]), dtype64=False)
    m.to(DEV)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(batch_from_packed(make_packed_graphs(2, seed=0)))          # CPU batch on a CUDA model


def test_full_size_properties_4096_graphs():
    """BASELINE config 2 size: properties that need no oracle run.
    (a) permuting the graphs of a batch permutes the eval outputs; (b) a batch equals the
    concatenation of its halves in eval mode; (c) the training loss is finite and gradients are
    non-zero for every live parameter."""
    c = cfg()
    packed = make_packed_graphs(4096, seed=1)
    _, _, m = make_pair(c, packed.in_degree_histogram(), dtype64=False)
    m.to(DEV).eval()
    g = torch.Generator().manual_seed(0)
    perm = torch.randperm(4096, generator=g)
    with torch.no_grad():
        y = m(batch_from_packed(packed).to(DEV))
        yp = m(batch_from_packed(packed, perm).to(DEV))
        ya = m(batch_from_packed(packed, torch.arange(0, 2048)).to(DEV))
        yb = m(batch_from_packed(packed, torch.arange(2048, 4096)).to(DEV))
    torch.testing.assert_close(yp, y[perm.to(DEV)], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(torch.cat([ya, yb]), y, rtol=1e-5, atol=1e-5)
    m.train()
    batch = batch_from_packed(packed).to(DEV)
    loss = M.mape(m(batch), batch.para.view(-1, 5))
    loss.backward()
    assert torch.isfinite(loss)
    mask = zero_grad_bias_mask(m)
    for (name, p) in m.named_parameters():
        assert torch.isfinite(p.grad).all(), name
    assert float(flat_grads(m)[mask].abs().max()) > 0


@pytest.mark.parametrize("precision,fwd_tol,grad_tol", [("tf32x3", 1e-5, 2e-3), ("simt", 1e-5, 2e-3), ("bf16", None, None)])
@pytest.mark.parametrize("name", ["default", "pre2_post2", "wide_small"])
def test_tensor_core_dense_stages(name, precision, fwd_tol, grad_tol):
    """Dense stages on tcgen05: TF32x3 must meet the fp32 parity bar; bf16 operands (the reference's
    own torch "medium" matmul precision) are held to 2 x the error measured on B200 (tests/helpers.py BF16_BARS;
    min/max selection and 1/std amplify operand rounding in the gradient)."""
    if precision == "bf16":
        fwd_tol, grad_tol = BF16_BARS[name]
    c = cfg(**CONFIGS[name])
    packed = make_packed_graphs(128, seed=6, num_para=c.num_para)
    batch = batch_from_packed(packed)
    o32, o64, m = make_pair(c, packed.in_degree_histogram())
    m.matmul_precision = precision
    m.to(DEV)
    for mod in (o32, o64, m):
        mod.eval()
    with torch.no_grad():
        y32, y64, y = o32(batch), o64(batch), m(batch.to(DEV))
        with bf16_operand_rounding(o64):        # floor of bf16 operand rounding on the reference arithmetic itself
            f_e = float((o64(batch) - y64).abs().max() / y64.abs().max())
    e_m, e_o, e_mo = _errs(y, y32, y64)
    print(f"{name}/{precision}: eval fwd cuda-vs-fp64 {e_m:.2e} (oracle32 {e_o:.2e})")
    assert e_m <= max(fwd_tol, 2 * e_o + 1e-6)
    for mod in (o64, m):
        mod.train()
    tgt = batch.para.view(-1, c.num_para)
    O.mape(o64(batch), tgt.double()).backward()
    M.mape(m(batch.to(DEV)), tgt.to(DEV)).backward()
    mask = zero_grad_bias_mask(m)
    g64, g = flat_grads(o64), flat_grads(m)
    rel = float((g - g64)[mask].norm() / g64[mask].norm())
    print(f"{name}/{precision}: grad rel-L2 {rel:.2e}")
    assert rel <= grad_tol
    if precision == "bf16":      # not worse than 2x the reference arithmetic with the same operands rounded to bf16
        with bf16_operand_rounding(o64):
            o64.zero_grad()
            yb = o64(batch)
            O.mape(yb, tgt.double()).backward()
            f_g = float((flat_grads(o64) - g64)[mask].norm() / g64[mask].norm())
        print(f"{name}/bf16 floor (fp64 oracle, bf16-rounded operands): eval fwd {f_e:.2e}, grad rel-L2 {f_g:.2e}")
        assert e_m <= 2 * f_e + 1e-6 and rel <= 2 * f_g


@pytest.mark.parametrize("train", [False, True])
def test_folded_destination_projection_matches_unfolded(train, monkeypatch):
    """csrc/fold.cu: moving Wdst x into the post_nns weights (default with self-loops) must reproduce the
    [Pdst | Psrc] formulation: forward within 1e-5 (max-normalised), gradients within the documented
    fp32 floor (global rel-L2 2e-3; measured far below)."""
    c = cfg()
    packed = make_packed_graphs(300, seed=11)
    batch = batch_from_packed(packed)
    _, _, m = make_pair(c, packed.in_degree_histogram(), dtype64=False)
    m.to(DEV).train(train)
    tgt = batch.para.view(-1, c.num_para).to(DEV)
    outs = {}
    for fold in (True, False):
        monkeypatch.setattr(M, "FOLD_DST", fold)
        assert m._fold_dst(m._matmul_mode()) == fold
        state = {k: v.clone() for k, v in m.state_dict().items()}
        m.zero_grad()
        y = m(batch.to(DEV))
        M.mape(y, tgt).backward()
        outs[fold] = (y.detach().double().cpu(), flat_grads(m).double().cpu())
        m.load_state_dict(state)           # same running statistics for the second pass
    (y1, g1), (y0, g0) = outs[True], outs[False]
    e = float((y1 - y0).abs().max() / y0.abs().max())
    mask = zero_grad_bias_mask(m)
    ge = float((g1 - g0)[mask].norm() / g0[mask].norm())
    print(f"folded vs unfolded (train={train}): forward {e:.2e}, gradient rel-L2 {ge:.2e}")
    assert e <= (2e-4 if train else 1e-5), e        # train mode: std-threshold flips (documented floor 1.4e-4)
    assert ge <= 2e-3, ge


def test_device_prefetcher_feeds_identical_batches():
    """data/loader.py: the side-stream H2D feed must deliver exactly the host batches, in order, and the training
    engine must produce the same losses as with synchronous `.to(device)` copies."""
    from gnnepcsaft_b200.data.loader import DevicePrefetcher
    from gnnepcsaft_b200.train.engine import TrainEngine
    c = cfg(propagation_depth=2)
    packed = make_packed_graphs(96, seed=5)
    host = [batch_from_packed(packed, torch.arange(i * 32, (i + 1) * 32)) for i in range(3)]
    got = list(DevicePrefetcher(iter(host), DEV, depth=2))
    assert len(got) == 3
    for h, d in zip(host, got):
        for k in ("x", "edge_index", "edge_attr", "batch", "para"):
            assert torch.equal(getattr(d, k).cpu(), getattr(h, k)), k
    losses = []
    for feed in ("sync", "prefetch"):
        _, _, m = make_pair(c, packed.in_degree_histogram(), dtype64=False)
        eng = TrainEngine.from_config(m.to(DEV), c)
        it = (b.to(DEV) for b in host) if feed == "sync" else DevicePrefetcher(iter(host), DEV)
        losses.append([float(eng.step(b)) for b in it])
    # the first loss is bit-identical (deterministic forward); later ones see parameters updated from gradients
    # whose split-K / bond-table sums use floating-point atomics (order not fixed)
    assert losses[0][0] == losses[1][0], losses
    torch.testing.assert_close(torch.tensor(losses[1]), torch.tensor(losses[0]), rtol=1e-4, atol=1e-6)


def test_two_pass_scatter_matches_atomic(monkeypatch):
    """GNN_B200_ATOMIC_SCATTER=0 (per-slot gradients + gnn_slot_sum_by_source, fixed summation order) against the
    default in-kernel atomic scatter: same gradients up to the summation order, and bit-identical from run to run
    in everything that does not go through the split-K / bond-table atomics."""
    c = cfg()
    packed = make_packed_graphs(200, seed=13)
    batch = batch_from_packed(packed).to(DEV)
    _, _, m = make_pair(c, packed.in_degree_histogram(), dtype64=False)
    m.to(DEV).eval()
    tgt = batch.para.view(-1, c.num_para)
    grads = {}
    for det in (False, True):
        monkeypatch.setattr(M, "DETERMINISTIC_SCATTER", det)
        m.zero_grad()
        M.mape(m(batch), tgt).backward()
        grads[det] = flat_grads(m).double().cpu()
    rel = float((grads[True] - grads[False]).norm() / grads[False].norm())
    print(f"two-pass vs atomic scatter: gradient rel-L2 {rel:.2e}")
    assert rel <= 1e-5, rel
