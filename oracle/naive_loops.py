"""ORACLE cross-check (test infrastructure): second, independent restatement of the
PNAPCSAFT forward as per-node / per-edge float64 numpy loops, written from the math in
SURVEY.md appendix A rather than from oracle/pna_oracle.py.  Small graphs only.

Consumes a reference-layout state_dict, eval-mode BatchNorm only.
Restates reference gnnepcsaft/train/models.py:105-135 and the PyG/OGB calls it makes.
"""

from __future__ import annotations

import math

import numpy as np


def _np(sd, key):
    return sd[key].detach().cpu().double().numpy()


def _linear(sd, prefix, v):
    return _np(sd, prefix + ".weight") @ v + _np(sd, prefix + ".bias")


def _seq(sd, prefix, v, n_layers):
    """Sequential(Linear, [ReLU, Linear]*(n-1)) with module indices 0,2,4,..."""
    v = _linear(sd, f"{prefix}.0", v)
    for i in range(1, n_layers):
        v = np.maximum(v, 0.0)
        v = _linear(sd, f"{prefix}.{2 * i}", v)
    return v


def _bn_eval(sd, prefix, v, eps=1e-5):
    return ((v - _np(sd, prefix + ".running_mean")) / np.sqrt(_np(sd, prefix + ".running_var") + eps)
            * _np(sd, prefix + ".weight") + _np(sd, prefix + ".bias"))


def forward_eval(sd, x, edge_index, edge_attr, batch, *, hidden, depth, pre_layers, post_layers,
                 num_mlp_layers, skip, self_loops, towers=2):
    x = np.asarray(x)
    edge_index = np.asarray(edge_index)
    edge_attr = np.asarray(edge_attr)
    n = x.shape[0]
    src = list(edge_index[0])
    dst = list(edge_index[1])
    codes = [tuple(r) for r in edge_attr]
    if self_loops:
        src += list(range(n))
        dst += list(range(n))
        codes += [(0, 0, 0)] * n
    f_in, f_out = hidden, hidden // towers

    h = np.zeros((n, hidden))
    for i in range(n):
        for k in range(9):
            h[i] += _np(sd, f"node_embed.atom_embedding_list.{k}.weight")[x[i, k]]
    bond = np.zeros((len(src), hidden))
    for e, c in enumerate(codes):
        for k in range(3):
            bond[e] += _np(sd, f"edge_embed.bond_embedding_list.{k}.weight")[c[k]]

    incoming = [[] for _ in range(n)]
    for e, d in enumerate(dst):
        incoming[d].append(e)

    for layer in range(depth):
        p = f"convs.{layer}"
        delta = float(_np(sd, f"{p}.aggr_module.avg_deg_log")[0])
        new = np.zeros((n, hidden))
        tower_out = np.zeros((n, towers * f_out))
        enc = [_linear(sd, f"{p}.edge_encoder", bond[e]) for e in range(len(src))]
        for i in range(n):
            d = len(incoming[i])
            amp = math.log(d + 1) / delta
            att = delta / math.log(max(d, 1) + 1)
            for t in range(towers):
                msgs = [_seq(sd, f"{p}.pre_nns.{t}",
                             np.concatenate([h[i], h[src[e]], enc[e]]), pre_layers)
                        for e in incoming[i]]
                if d == 0:
                    mean = mn = mx = std = np.zeros(f_in)
                else:
                    m = np.stack(msgs)
                    mean = m.sum(0) / d
                    mn = m.min(0)
                    mx = m.max(0)
                    var = (m * m).sum(0) / d - mean * mean
                    std = np.sqrt(np.maximum(var, 1e-5))
                    std = np.where(std <= math.sqrt(1e-5), 0.0, std)
                a = np.concatenate([mean, mn, mx, std])
                feat = np.concatenate([h[i], a, a * amp, a * att])
                tower_out[i, t * f_out:(t + 1) * f_out] = _seq(sd, f"{p}.post_nns.{t}", feat, post_layers)
            new[i] = _linear(sd, f"{p}.lin", tower_out[i])
        new = np.stack([_bn_eval(sd, f"batch_norms.{layer}.module", new[i]) for i in range(n)]) if n else new
        new = np.maximum(new, 0.0)
        h = new + h if skip else new

    if batch is None:
        pooled = h.sum(0, keepdims=True)
    else:
        batch = np.asarray(batch)
        pooled = np.zeros((int(batch.max()) + 1, hidden))
        for i in range(n):
            pooled[batch[i]] += h[i]

    out = []
    for row in pooled:
        v = row
        for i in range(num_mlp_layers):
            v = np.maximum(_bn_eval(sd, f"mlp.{4 * i + 1}", _linear(sd, f"mlp.{4 * i}", v)), 0.0)
        k = 4 * num_mlp_layers
        v = np.maximum(_bn_eval(sd, f"mlp.{k}.1", _linear(sd, f"mlp.{k}.0", v)), 0.0)
        v = np.maximum(_bn_eval(sd, f"mlp.{k}.5", _linear(sd, f"mlp.{k}.4", v)), 0.0)
        out.append(_linear(sd, f"mlp.{k}.8", v))
    return np.stack(out)
