"""Drop-in surface of the reference that surrounds the hot path (CPU; the C ABI is routed to oracle/abi_reference.py
by the `cpu_backend` fixture where kernels would run):

  * the `gnnepcsaft.*` import shim and the reference's call pattern `create_model(get_config(), calc_deg(name, workdir))`
    (reference train/train.py:128-129, train/utils.py:26-85, configs/default.py);
  * `PNApcsaftL.configure_optimizers` / `training_step` against torch.optim + the oracle (reference models.py:162-202);
  * `.pth` / `.ckpt` checkpoint round trips readable by the reference's `loadckp` (demo/utils.py:42-50) and
    `load_checkpoint` (train/utils.py:220-233), including the optimizer / schedule state of `TrainEngine`;
  * the `(Data, slices)` dataset reader, including a pickle that references torch_geometric classes;
  * malformed inputs raise like the reference's torch indexing does instead of being silently clamped.
"""
import os
import sys
import types

import pytest
import torch

from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
from gnnepcsaft_b200.train import models as M
from oracle import pna_oracle as O
from tests.helpers import cfg, make_pair


def _write_dataset(tmp_path, name="esper", num=20, seed=2, num_para=5):
    from gnnepcsaft_b200.data.dataset_io import processed_path, save_inmemory_dataset
    packed = make_packed_graphs(num, seed=seed, num_para=num_para)
    path = processed_path(name, str(tmp_path))
    save_inmemory_dataset(packed, path)
    return packed, path


def test_shim_reference_call_pattern(tmp_path, cpu_backend):
    import gnnepcsaft
    from gnnepcsaft.configs.configs_parallel import get_configs
    from gnnepcsaft.configs.default import get_config
    from gnnepcsaft.train.models import PNAPCSAFT, PNApcsaftL, PnaconvsParams, ReadoutMLPParams  # noqa: F401
    from gnnepcsaft.train.utils import calc_deg, create_model
    assert gnnepcsaft.__b200_shim__
    packed, _ = _write_dataset(tmp_path)
    config = get_config()
    # the fields of reference configs/default.py:6-46
    assert (config.model, config.hidden_dim, config.propagation_depth, config.batch_size, config.optimizer,
            config.warmup_steps, config.num_para, config.add_self_loops) == ("PNAL", 64, 6, 512, "adam", 100, 5, True)
    assert len(get_configs()) == 6 and get_configs()[4]["hidden_dim"] == 256 and get_configs()[3]["pre_layers"] == 2
    deg = calc_deg("esper", str(tmp_path))                       # reference train.py:128
    assert torch.equal(deg, packed.in_degree_histogram())
    assert torch.equal(deg, calc_deg(packed)) and torch.equal(deg, calc_deg(packed.to_list()))
    with pytest.raises(ValueError, match="ramirez or thermoml"):
        calc_deg("nope", str(tmp_path))
    model = create_model(config, deg)                            # reference train.py:129
    assert isinstance(model, PNApcsaftL) and isinstance(model.model, PNAPCSAFT)
    config.model = "PNA"
    assert isinstance(create_model(config, deg), PNAPCSAFT)
    config.model = "GAT"
    with pytest.raises(ValueError, match="Unsupported model"):
        create_model(config, deg)


def test_lightning_wrapper_training_step_and_optimizers(cpu_backend):
    c = cfg(propagation_depth=2, warmup_steps=3)
    packed = make_packed_graphs(10, seed=8)
    deg = packed.in_degree_histogram()
    o32, _, inner = make_pair(c, deg, dtype64=False)
    c.model = "PNAL"
    from gnnepcsaft_b200.train.utils import create_model
    lm = create_model(c, deg)
    lm.load_state_dict({"model." + k: v for k, v in o32.state_dict().items()}, strict=True)
    for hook in ("training_step", "validation_step", "test_step", "configure_optimizers", "log", "log_dict"):
        assert callable(getattr(lm, hook))
    cfg_opt = lm.configure_optimizers()
    opt, sch = cfg_opt["optimizer"], cfg_opt["lr_scheduler"]
    assert isinstance(opt, torch.optim.AdamW) and opt.defaults["amsgrad"] and opt.defaults["eps"] == 1e-5
    assert opt.defaults["weight_decay"] == c.weight_decay and opt.defaults["lr"] == c.learning_rate
    assert isinstance(sch["scheduler"], torch.optim.lr_scheduler.CosineAnnealingWarmRestarts)
    assert sch["scheduler"].T_0 == c.warmup_steps and (sch["interval"], sch["frequency"]) == ("step", 1)
    c.optimizer = "sgd"
    sgd = lm.configure_optimizers()["optimizer"]
    assert isinstance(sgd, torch.optim.SGD) and sgd.defaults["nesterov"] and sgd.defaults["momentum"] == c.momentum
    c.optimizer = "lion"
    with pytest.raises(ValueError, match="Unsupported optimizer"):
        lm.configure_optimizers()
    c.optimizer = "adam"

    # three Lightning-style steps (training_step -> zero_grad -> backward -> step -> scheduler.step) vs the oracle
    o32.train(), lm.train()
    ref_opt = torch.optim.AdamW(o32.parameters(), lr=c.learning_rate, weight_decay=c.weight_decay, amsgrad=True, eps=1e-5)
    ref_sch = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(ref_opt, c.warmup_steps)
    sched = sch["scheduler"]
    for it in range(3):
        batch = batch_from_packed(packed, torch.arange(it % 2 * 5, it % 2 * 5 + 5))
        loss = lm.training_step(batch, it)
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        ref_opt.zero_grad()
        ref_loss = O.mape(o32(batch), batch.para.view(-1, c.num_para))
        ref_loss.backward()
        ref_opt.step()
        ref_sch.step()
        assert abs(float(loss) - float(ref_loss)) <= 1e-4 * abs(float(ref_loss)) + 1e-6, it
        assert opt.param_groups[0]["lr"] == pytest.approx(ref_opt.param_groups[0]["lr"])
    lm.eval()                                                # Lightning runs validation in eval mode
    with torch.no_grad():
        out = lm.validation_step(packed.get(0), 0)
    assert out == {}
    lm.property_metrics = lambda para, graphs: {"mape_den": float(para.sum()) * 0.0}
    with torch.no_grad():
        assert lm.test_step(packed.get(0), 0) == {"mape_den": 0.0}


def test_checkpoint_roundtrips_through_the_shim(tmp_path, cpu_backend):
    from gnnepcsaft.train.utils import create_model
    from gnnepcsaft_b200.train.engine import TrainEngine
    from gnnepcsaft_b200.train.utils import load_checkpoint, loadckp, savemodel
    c = cfg(propagation_depth=2, warmup_steps=4)
    packed = make_packed_graphs(12, seed=4)
    deg = packed.in_degree_histogram()
    _, _, m = make_pair(c, deg, dtype64=False)
    eng = TrainEngine.from_config(m, c, distributed=False)
    for it in range(3):
        eng.step(batch_from_packed(packed, torch.arange(it * 4, it * 4 + 4)))
    pth = str(tmp_path / "train/checkpoints/last_checkpoint.pth")
    ckpt = str(tmp_path / "train/checkpoints/esper1-0-0.0100.ckpt")
    eng.save_checkpoint(pth)
    eng.save_checkpoint(ckpt)

    raw = torch.load(pth, weights_only=False)            # reference train/utils.py:109-119 layout
    assert set(raw) == {"model_state_dict", "optimizer_state_dict", "scaler_state_dict", "step"} and raw["step"] == 3
    assert list(raw["model_state_dict"]) == list(m.state_dict())
    # a torch AdamW over the same parameters accepts the optimizer state
    o = O.create_model(c, deg)
    o.load_state_dict(raw["model_state_dict"], strict=True)
    topt = torch.optim.AdamW(o.parameters(), lr=c.learning_rate, weight_decay=c.weight_decay, amsgrad=True, eps=1e-5)
    topt.load_state_dict(raw["optimizer_state_dict"])
    assert float(topt.state[next(iter(o.parameters()))]["step"]) == 3.0

    # reference demo/utils.py:42-50: PNAPCSAFT <- .pth, PNApcsaftL <- .ckpt["state_dict"] ("model." prefix)
    fresh = create_model(c, deg)
    loadckp(pth, fresh)
    for k, v in m.state_dict().items():
        assert torch.equal(fresh.state_dict()[k], v), k
    c.model = "PNAL"
    wrapped = create_model(c, deg)
    loadckp(ckpt, wrapped)
    for k, v in m.state_dict().items():
        assert torch.equal(wrapped.state_dict()["model." + k], v), k
    c.model = "PNA"

    # resume: a new engine continues bit-identically (same moments, same step count -> same cosine phase)
    m2 = create_model(c, deg)
    eng2 = TrainEngine.from_config(m2, c, distributed=False)
    c.change_opt = False
    _, initial_step = load_checkpoint(c, str(tmp_path), m2, eng2, None, "cpu")   # reference train/utils.py:220-233
    assert initial_step == 4 and int(eng2.step_count) == 3
    nxt = batch_from_packed(packed, torch.arange(0, 6))
    l1, l2 = eng.step(nxt), eng2.step(nxt)
    assert float(l1) == float(l2)
    for p1, p2 in zip(m.parameters(), m2.parameters()):
        assert torch.equal(p1, p2)
    m3 = create_model(c, deg)
    eng3 = TrainEngine.from_config(m3, c, distributed=False)
    assert eng3.load_checkpoint(ckpt) == 3 and int(eng3.step_count) == 3
    assert eng3.scheduler_state_dict()["T_cur"] == 3 % c.warmup_steps
    savemodel(m3, torch.optim.SGD(m3.parameters(), lr=0.1), None, str(tmp_path / "x.pth"), 7)
    assert torch.load(str(tmp_path / "x.pth"), weights_only=False)["step"] == 7


def test_dataset_reader_plain_and_pyg_pickle(tmp_path):
    from gnnepcsaft_b200.data.dataset_io import Esper, load_inmemory_dataset
    packed, path = _write_dataset(tmp_path, "esper", num=9, seed=6)
    back = load_inmemory_dataset(path)
    for k in ("x", "edge_index", "edge_attr", "para", "node_slices", "edge_slices"):
        assert torch.equal(getattr(back, k), getattr(packed, k)), k
    assert torch.equal(Esper(os.path.join(str(tmp_path), "data/esper2023")).x, packed.x)

    # the same content pickled the way PyG does it: Data._store (GlobalStorage) ._mapping, classes that only exist
    # while writing - reading must not need torch_geometric
    mods = {}
    for name in ("torch_geometric", "torch_geometric.data", "torch_geometric.data.data", "torch_geometric.data.storage"):
        mods[name] = types.ModuleType(name)
    class GlobalStorage:                                     # noqa: E306
        def __init__(self, mapping):
            self._mapping = mapping
            self._parent = None
    class Data:                                              # noqa: E306
        def __init__(self, mapping):
            self._store = GlobalStorage(mapping)
            self._edge_attr_cls = None
    GlobalStorage.__module__, GlobalStorage.__qualname__ = "torch_geometric.data.storage", "GlobalStorage"
    Data.__module__, Data.__qualname__ = "torch_geometric.data.data", "Data"
    mods["torch_geometric.data.storage"].GlobalStorage = GlobalStorage
    mods["torch_geometric.data.data"].Data = Data
    g = packed.num_graphs
    mapping = {"x": packed.x, "edge_index": packed.edge_index, "edge_attr": packed.edge_attr,
               "para": packed.para.reshape(-1), "InChI": [f"InChI=1S/{i}" for i in range(g)]}
    slices = {"x": packed.node_slices, "edge_index": packed.edge_slices, "edge_attr": packed.edge_slices,
              "para": torch.arange(g + 1) * 5, "InChI": torch.arange(g + 1)}
    pyg_path = str(tmp_path / "pyg_style.pt")
    sys.modules.update(mods)
    try:
        torch.save((Data(mapping), slices), pyg_path)
    finally:
        for name in mods:
            sys.modules.pop(name, None)
    assert "torch_geometric" not in sys.modules
    back = load_inmemory_dataset(pyg_path)
    for k in ("x", "edge_index", "edge_attr", "para", "node_slices", "edge_slices"):
        assert torch.equal(getattr(back, k), getattr(packed, k)), k
    with pytest.raises(ValueError):
        torch.save(({"x": packed.x}, {"x": packed.node_slices}), str(tmp_path / "bad.pt"))
        load_inmemory_dataset(str(tmp_path / "bad.pt"))


def test_malformed_inputs_raise_like_the_reference(cpu_backend):
    c = cfg(propagation_depth=1)
    packed = make_packed_graphs(4, seed=1)
    _, _, m = make_pair(c, packed.in_degree_histogram(), dtype64=False)
    m.eval()
    good = batch_from_packed(packed)
    with torch.no_grad():
        m(good)
        bad = batch_from_packed(packed)
        bad.x = bad.x.clone()
        bad.x[3, 0] = 119                                    # atomic-number index beyond the OGB table
        with pytest.raises(IndexError, match="atom feature"):
            m(bad)
        bad = batch_from_packed(packed)
        bad.batch = bad.batch.flip(0).contiguous()           # unsorted batch vector
        with pytest.raises(ValueError, match="not sorted"):
            m(bad)
        bad = batch_from_packed(packed)
        bad.batch = bad.batch.clone()
        bad.batch[-1] = 99
        with pytest.raises(IndexError, match="batch id"):
            m(bad)
    # training-mode BatchNorm on one row: torch raises, so does this path
    m.train()
    single = packed.get(0)
    single.x, single.edge_index, single.edge_attr = single.x[:1], single.edge_index[:, :0], single.edge_attr[:0]
    with pytest.raises(ValueError, match="more than 1 value per channel"):
        m(single)
