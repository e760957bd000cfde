"""Graph containers for the PNAPCSAFT hot path.

The reference hands `torch_geometric.data.Data` / `Batch` objects to
`PNAPCSAFT.forward(data)` (reference gnnepcsaft/train/models.py:105-116, built by
gnnepcsaft/data/graph.py:34-40 and the PyG DataLoader at gnnepcsaft/train/train.py:74-79).
`torch_geometric` is not a dependency of this package; `forward` only needs the four
attributes below, so any object exposing them (including a real PyG `Data`) works.
"""

from __future__ import annotations

from typing import List, Optional

import torch

# OGB integer feature space (ogb.utils.features, restated in SURVEY.md appendix A.1)
ATOM_FEATURE_DIMS = (119, 5, 12, 12, 10, 6, 6, 2, 2)
BOND_FEATURE_DIMS = (5, 6, 2)
NUM_BOND_CODES = BOND_FEATURE_DIMS[0] * BOND_FEATURE_DIMS[1] * BOND_FEATURE_DIMS[2]  # 60


class Data:
    """Minimal stand-in for `torch_geometric.data.Data`.

    x          int64 [N, 9]   OGB atom features
    edge_index int64 [2, E]   row 0 = source, row 1 = destination (aggregation index)
    edge_attr  int64 [E, 3]   OGB bond features
    batch      int64 [N] or None (single molecule, like PyG's `Data.batch` property)
    ptr        int64 [B+1] or None
    para       float [B*P] or None (training target, reference models.py:192)
    """

    def __init__(self, x=None, edge_index=None, edge_attr=None, batch=None, ptr=None,
                 para=None, num_graphs: Optional[int] = None, **extra):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.batch = batch
        self.ptr = ptr
        self.para = para
        self._num_graphs = num_graphs
        for k, v in extra.items():
            setattr(self, k, v)

    @property
    def num_nodes(self) -> int:
        return int(self.x.shape[0])

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.shape[1])

    @property
    def num_graphs(self) -> int:
        if self._num_graphs is not None:
            return self._num_graphs
        if self.ptr is not None:
            return int(self.ptr.numel()) - 1
        if self.batch is None:
            return 1
        return int(self.batch.max()) + 1

    def _apply(self, fn):
        out = Data.__new__(Data)
        for k, v in self.__dict__.items():
            out.__dict__[k] = fn(v) if isinstance(v, torch.Tensor) else v
        return out

    def to(self, *args, **kwargs):
        return self._apply(lambda t: t.to(*args, **kwargs))

    def pin_memory(self):
        return self._apply(lambda t: t.pin_memory())

    def __repr__(self):
        parts = [f"{k}={list(v.shape)}" for k, v in self.__dict__.items() if isinstance(v, torch.Tensor)]
        return f"Data({', '.join(parts)})"


Batch = Data


def collate(graphs: List[Data]) -> Data:
    """Host-side batch assembly with PyG `Batch.from_data_list` semantics.

    This is the host mirror used when graphs arrive as separate Python objects; the
    device path (`gnnepcsaft_b200.data.store.GraphStore.batch`) does the same work from a
    packed store in one kernel.  cat x / edge_attr along dim 0, cat edge_index along dim 1
    after adding the running node offset, batch = repeat_interleave(arange(B), n_g),
    ptr = [0, cumsum(n_g)], 1-D per-graph tensors (para) cat along dim 0.
    """
    counts = torch.tensor([g.x.shape[0] for g in graphs], dtype=torch.int64)
    ptr = torch.zeros(len(graphs) + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(counts, 0)
    x = torch.cat([g.x for g in graphs], 0)
    edge_attr = torch.cat([g.edge_attr for g in graphs], 0)
    edge_index = torch.cat([g.edge_index + int(ptr[i]) for i, g in enumerate(graphs)], 1)
    batch = torch.repeat_interleave(torch.arange(len(graphs), dtype=torch.int64), counts)
    para = None
    if all(getattr(g, "para", None) is not None for g in graphs):
        para = torch.cat([g.para.reshape(-1) for g in graphs], 0)
    return Data(x=x, edge_index=edge_index, edge_attr=edge_attr, batch=batch, ptr=ptr,
                para=para, num_graphs=len(graphs))
