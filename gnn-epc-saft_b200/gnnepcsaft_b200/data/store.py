"""Device-resident graph store + device collate.

Replaces the reference's `InMemoryDataset` + PyG `DataLoader(num_workers=os.cpu_count())`
(reference gnnepcsaft/data/graphdataset.py:48,137 and gnnepcsaft/train/train.py:74-79): the packed
store (concatenated tensors + slices, graph-local `edge_index`) is uploaded once; a batch is one
offset scan + one gather kernel that emits the PyG-layout int64 tensors bit-exactly.
"""

from __future__ import annotations

import torch

from .. import _lib
from .graph import Data


class GraphStore:
    def __init__(self, packed, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("gnnepcsaft_b200: GraphStore needs a CUDA device (no CPU path)")
        self.device = dev
        self.num_graphs = packed.num_graphs
        self.x = packed.x.to(dev).contiguous()
        self.edge_index = packed.edge_index.to(dev).contiguous()
        self.edge_attr = packed.edge_attr.to(dev).contiguous()
        self.para = packed.para.to(dev).contiguous()
        self.node_slices = packed.node_slices.to(dev).contiguous()
        self.edge_slices = packed.edge_slices.to(dev).contiguous()
        # host copies of the per-graph sizes: batch totals are computed without a device sync
        self._ncnt = (packed.node_slices[1:] - packed.node_slices[:-1]).clone()
        self._ecnt = (packed.edge_slices[1:] - packed.edge_slices[:-1]).clone()
        self._ws = None

    def batch(self, ids: torch.Tensor) -> Data:
        """ids: int64 graph indices, host or device tensor."""
        if ids.is_cuda:
            ids_dev = ids.contiguous()
            ids_host = ids.cpu()
        else:
            ids_host = ids.to(torch.int64)
            ids_dev = ids_host.to(self.device, non_blocking=True)
        nb = int(ids_host.numel())
        tot_n = int(self._ncnt[ids_host].sum())
        tot_e = int(self._ecnt[ids_host].sum())
        dev = self.device
        i64 = dict(dtype=torch.int64, device=dev)
        ptr = torch.empty(nb + 1, **i64)
        eptr = torch.empty(nb + 1, **i64)
        nbytes = _lib.query("gnn_collate_workspace_bytes", nb)
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        st = _lib.stream_ptr()
        _lib.call("gnn_collate_offsets", ids_dev, nb, self.node_slices, self.edge_slices, ptr, eptr, self._ws,
                  self._ws.numel(), st)
        x = torch.empty(tot_n, 9, **i64)
        edge_index = torch.empty(2, tot_e, **i64)
        edge_attr = torch.empty(tot_e, 3, **i64)
        batch = torch.empty(tot_n, **i64)
        _lib.call("gnn_collate_gather", ids_dev, nb, self.node_slices, self.edge_slices, self.x, self.edge_index,
                  int(self.edge_index.shape[1]), self.edge_attr, ptr, eptr, tot_n, tot_e, x, edge_index, edge_attr,
                  batch, st)
        para = self.para[ids_dev].reshape(-1)
        return Data(x=x, edge_index=edge_index, edge_attr=edge_attr, batch=batch, ptr=ptr, para=para, num_graphs=nb)
