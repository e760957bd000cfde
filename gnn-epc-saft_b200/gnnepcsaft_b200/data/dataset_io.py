"""Readers / writers for the reference's processed dataset files.

The reference stores every dataset as `torch.save((data, slices), processed_paths[0])` (gnnepcsaft/data/
graphdataset.py:137,253,318) and reloads it with `torch.load(..., weights_only=False)` (:48,221,284): `data` is ONE
PyG `Data` whose tensors are the concatenation over molecules (`edge_index` NOT incremented) and `slices[key]` are
int64 cumulative offsets of length `len(dataset) + 1`.  That is already the packed CSR-of-graphs the
device store wants, so no per-molecule Python objects are ever built here.

`torch_geometric` is not a dependency: the pickle is read with an unpickler that maps every `torch_geometric.*`
class to an inert attribute bag, then the tensor mapping is pulled out of it (PyG >= 2.0 keeps it in
`Data._store._mapping`; older versions kept tensors directly in `Data.__dict__`).  A plain `(dict, slices)` tuple -
what `save_inmemory_dataset` writes - loads without any of this.
"""

from __future__ import annotations

import os
import pickle
from typing import Dict, Optional, Tuple

import torch

from .synth import PackedGraphs


class _Bag:
    """Inert stand-in for any torch_geometric class met while unpickling."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):      # (dict, slots)
            for part in state:
                if isinstance(part, dict):
                    self.__dict__.update(part)
        elif isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state


_BAGS: Dict[Tuple[str, str], type] = {}


class _PygFreeUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in ("torch_geometric", "torch_sparse", "torch_scatter"):
            key = (module, name)
            if key not in _BAGS:
                _BAGS[key] = type(name, (_Bag,), {"__module__": module})
            return _BAGS[key]
        return super().find_class(module, name)


class _PygFreePickle:
    """`pickle_module` for torch.load."""
    __name__ = "gnnepcsaft_b200_pygfree_pickle"
    Unpickler = _PygFreeUnpickler
    load = staticmethod(lambda f, **kw: _PygFreeUnpickler(f, **kw).load())
    loads = staticmethod(pickle.loads)
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL


def _mapping_of(obj) -> dict:
    """Name -> value mapping of a (stubbed) PyG Data / storage object, or of a plain dict."""
    if isinstance(obj, dict):
        return obj
    d = getattr(obj, "__dict__", {})
    store = d.get("_store")
    if store is not None:
        m = _mapping_of(store)
        if m:
            return m
    if "_mapping" in d and isinstance(d["_mapping"], dict):
        return d["_mapping"]
    return {k: v for k, v in d.items() if not k.startswith("_")}


def load_data_slices(path: str):
    """`(mapping of concatenated tensors / lists, slices dict)` of a processed reference dataset file."""
    try:
        obj = torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        obj = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_PygFreePickle)
    if not (isinstance(obj, (tuple, list)) and len(obj) >= 2):
        raise ValueError(f"{path}: expected the reference's (data, slices) tuple")
    data, slices = _mapping_of(obj[0]), obj[1]
    if not isinstance(slices, dict):
        slices = _mapping_of(slices)
    for key in ("x", "edge_index", "edge_attr"):
        if key not in data or key not in slices:
            raise ValueError(f"{path}: '{key}' missing from the stored graph data")
    return data, slices


def packed_from_data_slices(data, slices, target: str = "para") -> PackedGraphs:
    """Packed store of the hot path's inputs (+ the per-molecule target vector `para`)."""
    node_slices = slices["x"].to(torch.int64)
    edge_slices = slices["edge_index"].to(torch.int64)
    g = int(node_slices.numel()) - 1
    if target in data and isinstance(data[target], torch.Tensor):
        para = data[target].to(torch.float32).reshape(g, -1) if g else data[target].to(torch.float32).reshape(0, 0)
    else:
        para = torch.zeros(g, 0, dtype=torch.float32)
    return PackedGraphs(x=data["x"].to(torch.int64).contiguous(),
                        edge_index=data["edge_index"].to(torch.int64).contiguous(),
                        edge_attr=data["edge_attr"].to(torch.int64).reshape(-1, 3).contiguous(), para=para,
                        node_slices=node_slices.contiguous(), edge_slices=edge_slices.contiguous())


def load_inmemory_dataset(path: str, target: str = "para") -> PackedGraphs:
    """Processed `InMemoryDataset` file of the reference -> `PackedGraphs` (host)."""
    return packed_from_data_slices(*load_data_slices(path), target=target)


def save_inmemory_dataset(packed: PackedGraphs, path: str, extra: Optional[dict] = None):
    """Write a packed store in the reference's `(data, slices)` layout with a plain dict in place of the PyG
    `Data` (loads with `weights_only=True`; PyG's `InMemoryDataset` accepts it through `Data.from_dict`)."""
    g = packed.num_graphs
    p = int(packed.para.shape[1]) if packed.para.dim() == 2 else 0
    data = {"x": packed.x, "edge_index": packed.edge_index, "edge_attr": packed.edge_attr,
            "para": packed.para.reshape(-1)}
    slices = {"x": packed.node_slices, "edge_index": packed.edge_slices, "edge_attr": packed.edge_slices,
              "para": torch.arange(g + 1, dtype=torch.int64) * p}
    if extra:
        data.update(extra.get("data", {}))
        slices.update(extra.get("slices", {}))
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save((data, slices), path)


# the reference's dataset classes, as readers of their processed files (graphdataset.py:140-318)
_PROCESSED = {"ramirez": ("data/ramirez2022", "ra_graph_data.pt"), "esper": ("data/esper2023", "es_graph_data.pt"),
              "thermoml": ("data/thermoml", "tml_graph_data.pt")}


def processed_path(dataset: str, workdir: str) -> str:
    try:
        sub, name = _PROCESSED[dataset]
    except KeyError:
        raise ValueError(f"dataset is either ramirez or thermoml, got >>> {dataset} <<< instead") from None
    return os.path.join(workdir, sub, "processed", name)


def Ramirez(root: str) -> PackedGraphs:          # noqa: N802 - the reference's class name
    """reference graphdataset.py:193-253 (para [3] per molecule)."""
    return load_inmemory_dataset(os.path.join(root, "processed", "ra_graph_data.pt"))


def Esper(root: str) -> PackedGraphs:            # noqa: N802
    """reference graphdataset.py:256-318 (para [5] per molecule)."""
    return load_inmemory_dataset(os.path.join(root, "processed", "es_graph_data.pt"))
