"""Synthetic OGB-featurised molecular graphs.

rdkit / ogb are not available, so graphs are generated directly in the integer feature
space `ogb.utils.mol.smiles2graph` emits (reference gnnepcsaft/data/graph.py:28-40):
x int64 [n,9], edge_index int64 [2,e] with the two directions of each bond adjacent,
edge_attr int64 [e,3] with both directions sharing one feature row.

The result is a *packed store* in the layout of the reference's `InMemoryDataset` files
(reference gnnepcsaft/data/graphdataset.py:48,137): every tensor is the
concatenation over graphs, `edge_index` is NOT incremented, and `slices` hold the segment
offsets.
"""

from __future__ import annotations

import dataclasses
from typing import List

import numpy as np
import torch

from .graph import ATOM_FEATURE_DIMS, BOND_FEATURE_DIMS, Data

MAX_VALENCE = 4


@dataclasses.dataclass
class PackedGraphs:
    """Concatenated graphs + slices (host, int64 / float32)."""

    x: torch.Tensor            # [sumN, 9] int64
    edge_index: torch.Tensor   # [2, sumE] int64, graph-local node ids
    edge_attr: torch.Tensor    # [sumE, 3] int64
    para: torch.Tensor         # [G, P] float32
    node_slices: torch.Tensor  # [G+1] int64
    edge_slices: torch.Tensor  # [G+1] int64

    @property
    def num_graphs(self) -> int:
        return int(self.node_slices.numel()) - 1

    def get(self, i: int) -> Data:
        """One molecule as the reference's `dataset[i]` would return it (batch=None)."""
        n0, n1 = int(self.node_slices[i]), int(self.node_slices[i + 1])
        e0, e1 = int(self.edge_slices[i]), int(self.edge_slices[i + 1])
        return Data(x=self.x[n0:n1], edge_index=self.edge_index[:, e0:e1],
                    edge_attr=self.edge_attr[e0:e1], para=self.para[i])

    def to_list(self) -> List[Data]:
        return [self.get(i) for i in range(self.num_graphs)]

    def in_degree_histogram(self) -> torch.Tensor:
        """`calc_deg` of the reference (train/utils.py:26-49): in-degree histogram, no self-loops."""
        sum_n = int(self.node_slices[-1])
        node_off = torch.repeat_interleave(self.node_slices[:-1],
                                           self.edge_slices[1:] - self.edge_slices[:-1])
        dst = self.edge_index[1] + node_off
        d = torch.bincount(dst, minlength=sum_n)
        return torch.bincount(d)


def _sizes(rng: np.random.Generator, num_graphs: int, skewed: bool) -> np.ndarray:
    if skewed:
        n = np.exp(rng.normal(np.log(20.0), 0.8, size=num_graphs))
        return np.clip(np.rint(n), 5, 150).astype(np.int64)
    n = rng.normal(25.0, 6.0, size=num_graphs)
    return np.clip(np.rint(n), 4, 60).astype(np.int64)


def make_packed_graphs(num_graphs: int, seed: int = 0, skewed: bool = False,
                       num_para: int = 5) -> PackedGraphs:
    """Random trees (max degree 4) + Poisson(1.2) ring closures, OGB integer features."""
    rng = np.random.default_rng(seed)
    sizes = _sizes(rng, num_graphs, skewed)
    n_rings = rng.poisson(1.2, size=num_graphs)

    src_all: List[np.ndarray] = []
    n_bonds = np.zeros(num_graphs, dtype=np.int64)
    for g in range(num_graphs):
        n = int(sizes[g])
        u = rng.random(n + 2 * int(n_rings[g]))
        deg = [0] * n
        avail = [0]                      # nodes with deg < MAX_VALENCE, among those placed
        bonds = []
        bonded = set()
        for k in range(1, n):
            pos = int(u[k] * len(avail))
            j = avail[pos]
            bonds.append((j, k))
            bonded.add((j, k))
            deg[j] += 1
            deg[k] = 1
            if deg[j] >= MAX_VALENCE:
                avail[pos] = avail[-1]
                avail.pop()
            avail.append(k)
        for r in range(int(n_rings[g])):
            a = int(u[n + 2 * r] * n)
            b = int(u[n + 2 * r + 1] * n)
            if a == b:
                continue
            lo, hi = (a, b) if a < b else (b, a)
            if deg[a] >= MAX_VALENCE or deg[b] >= MAX_VALENCE or (lo, hi) in bonded:
                continue
            bonds.append((a, b))
            bonded.add((lo, hi))
            deg[a] += 1
            deg[b] += 1
        arr = np.asarray(bonds, dtype=np.int64).reshape(-1, 2)
        # both directions adjacent: (a,b),(b,a)
        e = np.empty((2, 2 * arr.shape[0]), dtype=np.int64)
        e[0, 0::2] = arr[:, 0]
        e[1, 0::2] = arr[:, 1]
        e[0, 1::2] = arr[:, 1]
        e[1, 1::2] = arr[:, 0]
        src_all.append(e)
        n_bonds[g] = arr.shape[0]

    sum_n = int(sizes.sum())
    sum_b = int(n_bonds.sum())
    x = np.stack([rng.integers(0, d, size=sum_n) for d in ATOM_FEATURE_DIMS], axis=1).astype(np.int64)
    bond_feat = np.stack([rng.integers(0, d, size=sum_b) for d in BOND_FEATURE_DIMS], axis=1).astype(np.int64)
    edge_attr = np.repeat(bond_feat, 2, axis=0)
    edge_index = np.concatenate(src_all, axis=1) if src_all else np.zeros((2, 0), np.int64)
    para = rng.uniform(0.5, 5.0, size=(num_graphs, num_para)).astype(np.float32)

    node_slices = np.zeros(num_graphs + 1, dtype=np.int64)
    node_slices[1:] = np.cumsum(sizes)
    edge_slices = np.zeros(num_graphs + 1, dtype=np.int64)
    edge_slices[1:] = np.cumsum(2 * n_bonds)
    return PackedGraphs(
        x=torch.from_numpy(x), edge_index=torch.from_numpy(edge_index),
        edge_attr=torch.from_numpy(edge_attr), para=torch.from_numpy(para),
        node_slices=torch.from_numpy(node_slices), edge_slices=torch.from_numpy(edge_slices))


def batch_from_packed(p: PackedGraphs, ids=None) -> Data:
    """Host collate of a packed store with PyG semantics (vectorised)."""
    if ids is None:
        ids = torch.arange(p.num_graphs, dtype=torch.int64)
    ids = torch.as_tensor(ids, dtype=torch.int64)
    n_cnt = p.node_slices[ids + 1] - p.node_slices[ids]
    e_cnt = p.edge_slices[ids + 1] - p.edge_slices[ids]
    ptr = torch.zeros(ids.numel() + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(n_cnt, 0)
    eptr = torch.zeros(ids.numel() + 1, dtype=torch.int64)
    eptr[1:] = torch.cumsum(e_cnt, 0)
    g_of_node = torch.repeat_interleave(torch.arange(ids.numel()), n_cnt)
    g_of_edge = torch.repeat_interleave(torch.arange(ids.numel()), e_cnt)
    node_src = p.node_slices[ids][g_of_node] + (torch.arange(int(ptr[-1])) - ptr[:-1][g_of_node])
    edge_src = p.edge_slices[ids][g_of_edge] + (torch.arange(int(eptr[-1])) - eptr[:-1][g_of_edge])
    x = p.x[node_src]
    edge_attr = p.edge_attr[edge_src]
    edge_index = p.edge_index[:, edge_src] + ptr[:-1][g_of_edge]
    return Data(x=x, edge_index=edge_index, edge_attr=edge_attr, batch=g_of_node, ptr=ptr,
                para=p.para[ids].reshape(-1), num_graphs=int(ids.numel()))
