"""ORACLE (test infrastructure, not product code) — CPU restatement of `PNAPCSAFT`.

PARITY UNPINNED: the arithmetic of this path lives in un-vendored third-party packages
(torch_geometric ~2.6, ogb ~1.3.6; reference requirements.txt:1-25 pins nothing) that are
not installable here, and the reference ships no tests, golden tensors or checkpoints
(SURVEY.md §4, §8c).  This file restates their published algorithms with plain torch ops
in the op order the reference executes, and is cross-checked against an independent
per-node float64 loop restatement (oracle/naive_loops.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.

Restated call sites (reference gnnepcsaft/train/models.py):
  :65  AtomEncoder  -> OGB graphproppred/mol_encoder.py     (class AtomEncoder below)
  :66  BondEncoder  -> OGB graphproppred/mol_encoder.py     (class BondEncoder)
  :69-80 PNAConv    -> PyG nn/conv/pna_conv.py              (class PNAConv)
         aggr       -> PyG nn/aggr/scaler.py, multi.py, basic.py, utils/_scatter.py
                                                             (class DegreeScalerAggregation)
  :82  BatchNorm    -> PyG nn/norm/batch_norm.py             (class BatchNorm, `.module`)
  :118-121 add_self_loops -> PyG utils/loop.py               (add_self_loops)
  :133 global_add_pool    -> PyG nn/pool/glob.py             (global_add_pool)
  :84-103 mlp head, :105-135 forward                          (class PNAPCSAFT)
  :194 torchmetrics MAPE                                       (mape)
state_dict keys equal the reference model's.
"""

from __future__ import annotations

import dataclasses
import math

import torch
import torch.nn.functional as F
from torch import nn

ATOM_FEATURE_DIMS = (119, 5, 12, 12, 10, 6, 6, 2, 2)
BOND_FEATURE_DIMS = (5, 6, 2)
TOWERS = 2


@dataclasses.dataclass
class PnaconvsParams:
    propagation_depth: int
    pre_layers: int
    post_layers: int
    deg: torch.Tensor
    dropout: float = 0.0
    skip_connections: bool = False
    self_loops: bool = False


@dataclasses.dataclass
class ReadoutMLPParams:
    num_mlp_layers: int
    num_para: int
    dropout: float = 0.0


class AtomEncoder(nn.Module):
    def __init__(self, emb_dim: int):
        super().__init__()
        self.atom_embedding_list = nn.ModuleList()
        for dim in ATOM_FEATURE_DIMS:
            emb = nn.Embedding(dim, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            self.atom_embedding_list.append(emb)

    def forward(self, x):
        out = 0
        for i in range(x.shape[1]):
            out = out + self.atom_embedding_list[i](x[:, i])
        return out


class BondEncoder(nn.Module):
    def __init__(self, emb_dim: int):
        super().__init__()
        self.bond_embedding_list = nn.ModuleList()
        for dim in BOND_FEATURE_DIMS:
            emb = nn.Embedding(dim, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            self.bond_embedding_list.append(emb)

    def forward(self, edge_attr):
        out = 0
        for i in range(edge_attr.shape[1]):
            out = out + self.bond_embedding_list[i](edge_attr[:, i])
        return out


def add_self_loops(edge_index, edge_attr, fill_value, num_nodes):
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    loop = loop.view(1, -1).repeat(2, 1)
    loop_attr = edge_attr.new_full((num_nodes,) + tuple(edge_attr.shape[1:]), fill_value)
    return torch.cat([edge_index, loop], 1), torch.cat([edge_attr, loop_attr], 0)


def _scatter_sum(src, index, dim_size):
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return src.new_zeros((dim_size,) + tuple(src.shape[1:])).scatter_add_(0, idx, src)


def _scatter_mean(src, index, dim_size):
    count = src.new_zeros(dim_size).scatter_add_(0, index, src.new_ones(src.shape[0]))
    count = count.clamp(min=1)
    out = _scatter_sum(src, index, dim_size)
    return out / count.view(-1, *([1] * (src.dim() - 1)))


def _scatter_minmax(src, index, dim_size, which):
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    return out.scatter_reduce_(0, idx, src, reduce=which, include_self=False)


class DegreeScalerAggregation(nn.Module):
    """aggr = [mean, min, max, std], scaler = [identity, amplification, attenuation]."""

    def __init__(self, deg: torch.Tensor):
        super().__init__()
        deg = deg.to(torch.float)
        num_nodes = int(deg.sum())
        bin_degrees = torch.arange(deg.numel())
        avg_lin = float((bin_degrees * deg).sum()) / num_nodes
        avg_log = float(((bin_degrees + 1).log() * deg).sum()) / num_nodes
        self.register_buffer("avg_deg_lin", torch.tensor([avg_lin], dtype=torch.float))
        self.register_buffer("avg_deg_log", torch.tensor([avg_log], dtype=torch.float))

    def forward(self, m, index, dim_size):
        mean = _scatter_mean(m, index, dim_size)
        mn = _scatter_minmax(m, index, dim_size, "amin")
        mx = _scatter_minmax(m, index, dim_size, "amax")
        mean_2 = _scatter_mean(m * m, index, dim_size)
        var = mean_2 - mean * mean
        std = var.clamp(min=1e-5).sqrt()
        std = std.masked_fill(std <= math.sqrt(1e-5), 0.0)
        out = torch.cat([mean, mn, mx, std], dim=-1)

        deg = m.new_zeros(dim_size).scatter_add_(0, index, m.new_ones(index.numel()))
        deg = deg.view(-1, 1, 1)
        amp = out * (torch.log(deg + 1) / self.avg_deg_log)
        att = out * (self.avg_deg_log / torch.log(deg.clamp(min=1) + 1))
        return torch.cat([out, amp, att], dim=-1)


class PNAConv(nn.Module):
    def __init__(self, in_channels, out_channels, deg, edge_dim, towers, pre_layers, post_layers):
        super().__init__()
        self.towers = towers
        self.F_in = in_channels                      # divide_input=False
        self.F_out = out_channels // towers
        self.aggr_module = DegreeScalerAggregation(deg)
        self.edge_encoder = nn.Linear(edge_dim, self.F_in)
        self.pre_nns = nn.ModuleList()
        self.post_nns = nn.ModuleList()
        for _ in range(towers):
            mods = [nn.Linear(3 * self.F_in, self.F_in)]
            for _ in range(pre_layers - 1):
                mods += [nn.ReLU(), nn.Linear(self.F_in, self.F_in)]
            self.pre_nns.append(nn.Sequential(*mods))
            mods = [nn.Linear(13 * self.F_in, self.F_out)]
            for _ in range(post_layers - 1):
                mods += [nn.ReLU(), nn.Linear(self.F_out, self.F_out)]
            self.post_nns.append(nn.Sequential(*mods))
        self.lin = nn.Linear(out_channels, out_channels)

    def messages(self, x, edge_index, edge_attr):
        n = x.shape[0]
        xr = x.view(n, 1, self.F_in).repeat(1, self.towers, 1)
        e = self.edge_encoder(edge_attr)
        x_i = xr.index_select(0, edge_index[1])
        x_j = xr.index_select(0, edge_index[0])
        e = e.view(-1, 1, self.F_in).repeat(1, self.towers, 1)
        h = torch.cat([x_i, x_j, e], dim=-1)
        return xr, torch.stack([nn_(h[:, t]) for t, nn_ in enumerate(self.pre_nns)], dim=1)

    def forward(self, x, edge_index, edge_attr):
        n = x.shape[0]
        xr, m = self.messages(x, edge_index, edge_attr)
        s = self.aggr_module(m, edge_index[1], n)
        out = torch.cat([xr, s], dim=-1)
        out = torch.cat([nn_(out[:, t]) for t, nn_ in enumerate(self.post_nns)], dim=1)
        return self.lin(out)


class BatchNorm(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.module = nn.BatchNorm1d(in_channels)

    def forward(self, x):
        return self.module(x)


def global_add_pool(x, batch, size=None):
    if batch is None:
        return x.sum(dim=-2, keepdim=True)
    size = int(batch.max()) + 1 if size is None else size
    return _scatter_sum(x, batch, size)


def mape(pred, target):
    return torch.sum((pred - target).abs() / target.abs().clamp(min=1.17e-06)) / target.numel()


class PNAPCSAFT(nn.Module):
    def __init__(self, hidden_dim: int, pna_params: PnaconvsParams, mlp_params: ReadoutMLPParams):
        super().__init__()
        self.pna_params = pna_params
        self.convs = nn.ModuleList()
        self.batch_norms = nn.ModuleList()
        self.node_embed = AtomEncoder(hidden_dim)
        self.edge_embed = BondEncoder(hidden_dim)
        for _ in range(pna_params.propagation_depth):
            self.convs.append(PNAConv(hidden_dim, hidden_dim, pna_params.deg, hidden_dim, TOWERS,
                                      pna_params.pre_layers, pna_params.post_layers))
            self.batch_norms.append(BatchNorm(hidden_dim))
        self.mlp = nn.Sequential()
        for _ in range(mlp_params.num_mlp_layers):
            self.mlp.append(nn.Linear(hidden_dim, hidden_dim))
            self.mlp.append(nn.BatchNorm1d(hidden_dim))
            self.mlp.append(nn.ReLU())
            self.mlp.append(nn.Dropout(p=mlp_params.dropout))
        self.mlp.append(nn.Sequential(
            nn.Linear(hidden_dim, hidden_dim // 2), nn.BatchNorm1d(hidden_dim // 2), nn.ReLU(),
            nn.Dropout(p=mlp_params.dropout),
            nn.Linear(hidden_dim // 2, hidden_dim // 4), nn.BatchNorm1d(hidden_dim // 4), nn.ReLU(),
            nn.Dropout(p=mlp_params.dropout),
            nn.Linear(hidden_dim // 4, mlp_params.num_para)))

    def forward(self, data):
        x, edge_index, edge_attr, batch = data.x, data.edge_index, data.edge_attr, data.batch
        if self.pna_params.self_loops:
            edge_index, edge_attr = add_self_loops(edge_index, edge_attr, 0, x.size(0))
        x = self.node_embed(x)
        edge_attr = self.edge_embed(edge_attr)
        for conv, batch_norm in zip(self.convs, self.batch_norms):
            if self.pna_params.skip_connections:
                x_previous = x
            x = F.relu(batch_norm(conv(x, edge_index, edge_attr)))
            x = F.dropout(x, p=self.pna_params.dropout, training=self.training)
            if self.pna_params.skip_connections:
                x = x + x_previous
        x = global_add_pool(x, batch)
        return self.mlp(x)


def create_model(config, deg):
    """reference train/utils.py:52-85 for config.model == 'PNA'."""
    pna = PnaconvsParams(config.propagation_depth, config.pre_layers, config.post_layers, deg,
                         skip_connections=config.skip_connections, self_loops=config.add_self_loops)
    mlp = ReadoutMLPParams(config.num_mlp_layers, config.num_para, dropout=config.dropout_rate)
    return PNAPCSAFT(config.hidden_dim, pna, mlp)


def adamw_amsgrad_reference(model, lr=1e-3, weight_decay=1e-2):
    """reference models.py:164-170."""
    return torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=weight_decay,
                             amsgrad=True, eps=1e-5)
