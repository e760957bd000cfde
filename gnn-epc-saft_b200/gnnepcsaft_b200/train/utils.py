"""`create_model` / `calc_deg` of the reference's gnnepcsaft/train/utils.py:26-85."""

from __future__ import annotations

import types

import torch

from . import models


def calc_deg(dataset) -> torch.Tensor:
    """In-degree histogram of the training set WITHOUT self-loops (reference train/utils.py:26-49).

    The reference takes a dataset *name* and loads the PyG InMemoryDataset from disk; here any
    iterable of graphs (objects with `.edge_index`, `.x`) or a `PackedGraphs` store is accepted.
    """
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
    """reference gnnepcsaft/configs/default.py:6-46 (only the fields the hot path reads)."""
    return types.SimpleNamespace(
        optimizer="adam", learning_rate=1.0e-3, batch_size=512, warmup_steps=100, momentum=0.9,
        weight_decay=1e-2, model_name="esper1", model="PNAL", propagation_depth=6, hidden_dim=64,
        pre_layers=1, post_layers=1, dropout_rate=0.0, num_mlp_layers=1, num_para=5,
        skip_connections=True, add_self_loops=True)


def wide_config():
    """reference gnnepcsaft/configs/configs_parallel.py:43-51 (entry #5, the wide/deep variant)."""
    cfg = default_config()
    cfg.hidden_dim, cfg.propagation_depth, cfg.post_layers, cfg.add_self_loops = 256, 6, 2, False
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
