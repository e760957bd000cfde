"""`create_model` / `calc_deg` of the reference's gnnepcsaft/train/utils.py:26-85."""

from __future__ import annotations

import os
from typing import Optional

import torch

from . import models


def calc_deg(dataset, workdir: Optional[str] = None) -> torch.Tensor:
    """In-degree histogram of the training set WITHOUT self-loops (reference train/utils.py:26-49).

    Reference call pattern: `calc_deg("esper" | "ramirez", workdir)` - the processed `InMemoryDataset` file under
    `workdir/data/...` is read as a packed store (data/dataset_io.py; no PyG needed).  Also accepted: a
    `PackedGraphs` store (anything with `.in_degree_histogram()`) or any iterable of graphs with `.edge_index`, `.x`.
    """
    if isinstance(dataset, str):
        if workdir is None:
            raise TypeError("calc_deg(dataset: str, workdir: str): workdir is required with a dataset name")
        if dataset not in ("ramirez", "esper"):
            raise ValueError(f"dataset is either ramirez or thermoml, got >>> {dataset} <<< instead")
        from ..data.dataset_io import load_inmemory_dataset, processed_path
        dataset = load_inmemory_dataset(processed_path(dataset, workdir))
    if hasattr(dataset, "in_degree_histogram"):
        return dataset.in_degree_histogram()
    max_degree = -1
    degs = []
    for data in dataset:
        n = int(data.x.shape[0])
        d = torch.bincount(data.edge_index[1].cpu(), minlength=n)
        degs.append(d)
        max_degree = max(max_degree, int(d.max()) if n else 0)
    deg = torch.zeros(max_degree + 1, dtype=torch.long)
    for d in degs:
        deg += torch.bincount(d, minlength=deg.numel())
    return deg


def default_config():
    """reference gnnepcsaft/configs/default.py:6-46."""
    from ..configs.default import get_config
    return get_config()


def wide_config():
    """reference gnnepcsaft/configs/configs_parallel.py:43-51 (entry #5, the wide/deep variant)."""
    from ..configs.configs_parallel import get_configs
    cfg = default_config()
    for k, v in get_configs()[4].items():
        setattr(cfg, k, v)
    return cfg


def create_model(config, deg: torch.Tensor) -> torch.nn.Module:
    """Creates a model, as specified by the config (reference train/utils.py:52-85).

    `config` is any attribute-style object (ml_collections.ConfigDict, SimpleNamespace, ...)."""
    pna_params = models.PnaconvsParams(
        propagation_depth=config.propagation_depth, pre_layers=config.pre_layers,
        post_layers=config.post_layers, deg=deg, skip_connections=config.skip_connections,
        self_loops=config.add_self_loops)
    mlp_params = models.ReadoutMLPParams(num_mlp_layers=config.num_mlp_layers, num_para=config.num_para,
                                         dropout=config.dropout_rate)
    if config.model == "PNA":
        return models.PNAPCSAFT(hidden_dim=config.hidden_dim, pna_params=pna_params, mlp_params=mlp_params)
    if config.model == "PNAL":
        return models.PNApcsaftL(pna_params=pna_params, mlp_params=mlp_params, config=config)
    raise ValueError(f"Unsupported model: {config.model}.")


def create_optimizer(config, params):
    """Creates an optimizer, as specified by the config (reference train/utils.py:88-106)."""
    if config.optimizer == "adam":
        return torch.optim.AdamW(params, lr=config.learning_rate, weight_decay=config.weight_decay, amsgrad=True,
                                 eps=1e-5)
    if config.optimizer == "sgd":
        return torch.optim.SGD(params, lr=config.learning_rate, momentum=config.momentum,
                               weight_decay=config.weight_decay, nesterov=True)
    raise ValueError(f"Unsupported optimizer: {config.optimizer}.")


class _NoScaler:
    """Stands where the reference passes a `torch.amp.GradScaler` (config.amp is False everywhere)."""

    def state_dict(self):
        return {}

    def load_state_dict(self, state):
        return None


def savemodel(model, optimizer, scaler, path, step):
    """To checkpoint model during training (reference train/utils.py:109-119): the legacy `.pth` layout
    {"model_state_dict", "optimizer_state_dict", "scaler_state_dict", "step"} read by evaluate_ensemble.py:74-76 and
    demo/utils.py:42-50.  `optimizer` may be a torch optimizer or a `TrainEngine` (its AdamW-style state)."""
    opt_state = optimizer.optimizer_state_dict() if hasattr(optimizer, "optimizer_state_dict") else optimizer.state_dict()
    scaler = scaler if scaler is not None else _NoScaler()
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save({"model_state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
                "optimizer_state_dict": opt_state, "scaler_state_dict": scaler.state_dict(), "step": step}, path)


def save_lightning_checkpoint(model, path, engine=None, epoch: int = 0, global_step: int = 0):
    """`.ckpt` in the layout Lightning's ModelCheckpoint writes (reference train.py:86-107) as far as the reference
    reads it back: `checkpoint["state_dict"]` with the `model.` prefix of `PNApcsaftL` (demo/utils.py:42-50,
    train.py:170-175), plus `optimizer_states` / `lr_schedulers` / `epoch` / `global_step` for `trainer.fit(ckpt_path=)`."""
    sd = model.state_dict()
    if not isinstance(model, models.PNApcsaftL):
        sd = {"model." + k: v for k, v in sd.items()}
    ckpt = {"state_dict": {k: v.detach().cpu() for k, v in sd.items()}, "epoch": int(epoch),
            "global_step": int(global_step), "pytorch-lightning_version": "2.0.0", "loops": {}, "callbacks": {},
            "optimizer_states": [], "lr_schedulers": []}
    if engine is not None:
        ckpt["optimizer_states"] = [engine.optimizer_state_dict()]
        ckpt["lr_schedulers"] = [engine.scheduler_state_dict()]
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(ckpt, path)


def load_checkpoint(config, workdir, model, optimizer, scaler, device):
    "Loads saved model checkpoints (reference train/utils.py:220-233)."
    ckp_path = os.path.join(workdir, "train/checkpoints/last_checkpoint.pth")
    initial_step = 1
    if os.path.exists(ckp_path):
        checkpoint = torch.load(ckp_path, map_location=device, weights_only=False)
        model.load_state_dict(checkpoint["model_state_dict"])
        if not getattr(config, "change_opt", False) and optimizer is not None:
            if hasattr(optimizer, "load_optimizer_state_dict"):
                optimizer.load_optimizer_state_dict(checkpoint["optimizer_state_dict"])
            else:
                optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
        if scaler is not None:
            scaler.load_state_dict(checkpoint["scaler_state_dict"])
        initial_step = int(checkpoint["step"]) + 1
        del checkpoint
    return ckp_path, initial_step


def loadckp(ckp_path: str, model):
    """Loads save checkpoint (reference demo/utils.py:42-50): `.pth` for PNAPCSAFT, Lightning `.ckpt` for PNApcsaftL."""
    if os.path.exists(ckp_path):
        state = "model_state_dict" if isinstance(model, models.PNAPCSAFT) else "state_dict"
        checkpoint = torch.load(ckp_path, map_location=torch.device("cpu"), weights_only=False)
        model.load_state_dict(checkpoint[state])
        del checkpoint
