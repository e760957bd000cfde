"""Ensemble batched inference: K checkpoints of one architecture over a whole batch of molecules in one pass.

Replaces the loops of the reference's ensemble evaluation (gnnepcsaft/evaluations/evaluate_ensemble.py:128-208: for
every molecule of the loader, for every model of `model_dict`, one forward, then `torch.concat(pred).mean(0)`) and of
`demo/utils.py:20-40`: the batch's graph structure is built ONCE, the K forwards reuse it, predictions come back as
`[M, K, P]` (molecule, model, parameter) and the ensemble mean as `[M, P]`, reduced on the device.

Precision policy: the reference evaluates with `MODEL_DTYPE = torch.float64` models (evaluate_ensemble.py:68,
demo/utils.py:26).  The kernels of this package are fp32; `dtype=torch.float64` returns the fp32 result upcast, which
is within 1e-5 (max-normalised) of the float64 reference arithmetic (tests/test_ensemble_*.py, the same bar as the
single-model parity tests).
"""

from __future__ import annotations

from typing import Iterable, Optional, Sequence

import torch

from . import models as M
from .engine import TrainEngine
from .models import GraphStructure, PNAPCSAFT, _forward_impl


class EnsemblePredictor:
    def __init__(self, nets: Sequence[torch.nn.Module]):
        nets = [n.model if isinstance(n, M.PNApcsaftL) else n for n in nets]
        if not nets:
            raise ValueError("EnsemblePredictor needs at least one model")
        ref = nets[0]
        for n in nets:
            if not isinstance(n, PNAPCSAFT):
                raise TypeError("EnsemblePredictor takes PNAPCSAFT / PNApcsaftL models")
            if (n.hidden_dim, n.pna_params.propagation_depth, n.pna_params.self_loops, n.mlp_params.num_para) != \
                    (ref.hidden_dim, ref.pna_params.propagation_depth, ref.pna_params.self_loops, ref.mlp_params.num_para):
                raise ValueError("the models of an ensemble share one architecture")
            n.eval()
        self.nets = list(nets)
        self.device = ref.arena().device
        M._require_cuda(ref.arena(), "the models")
        self._engine = TrainEngine(ref, distributed=False)       # shape-bucket / CUDA-graph cache for the store path

    @classmethod
    def from_state_dicts(cls, config, deg: torch.Tensor, state_dicts: Iterable[dict], device="cuda"):
        """`state_dict`s as the reference's checkpoints hold them (`checkpoint["model_state_dict"]`, or a Lightning
        `state_dict` whose keys carry the `model.` prefix)."""
        from .utils import create_model
        nets = []
        for sd in state_dicts:
            sd = {k[len("model."):] if k.startswith("model.") else k: v.float() for k, v in sd.items()}
            cfg_model = getattr(config, "model", "PNA")
            config.model = "PNA"
            net = create_model(config, deg)
            config.model = cfg_model
            net.load_state_dict(sd, strict=True)
            nets.append(net.to(device))
        return cls(nets)

    @classmethod
    def from_checkpoints(cls, config, deg: torch.Tensor, paths: Iterable[str], device="cuda"):
        """`.pth` (reference train/utils.py:109-119) or Lightning `.ckpt` files."""
        sds = []
        for p in paths:
            ckpt = torch.load(p, map_location="cpu", weights_only=False)
            sds.append(ckpt["model_state_dict"] if "model_state_dict" in ckpt else ckpt["state_dict"])
        return cls.from_state_dicts(config, deg, sds, device)

    def __len__(self):
        return len(self.nets)

    # ---- one batch ------------------------------------------------------------------------------
    def _fresh_plans(self, batch):
        """The captured ensemble graph holds no weight-preparation kernels (see TrainEngine._fresh_plan)."""
        kind = "tiles" if (M.TILE_INFER and getattr(batch, "tileable", False)) else "layers"
        if any(n._matmul_mode() >= 0 and n._infer_plan(self.device, kind).filling for n in self.nets):
            self._predict_all(batch)

    def _predict_all(self, batch) -> torch.Tensor:
        """[rows, K, P]: rows = molecules (+ the padding graph of a compact batch)."""
        compact = getattr(batch, "is_compact", False)
        g = batch
        if not compact:
            bv = getattr(batch, "batch", None)
            nb = int(batch.num_graphs) if bv is not None else 1
            g = GraphStructure(batch.edge_index.contiguous(), batch.edge_attr.contiguous(),
                               bv.contiguous() if bv is not None else None, int(batch.x.shape[0]), nb,
                               self.nets[0].pna_params.self_loops)
        outs = []
        for net in self.nets:
            net.arena()
            if compact:
                out, _ = _forward_impl(net, None, None, None, None, batch.b, save=False, graph=batch, training=False)
            else:
                out, _ = _forward_impl(net, batch.x, batch.edge_index, batch.edge_attr, bv, nb, save=False, graph=g,
                                       training=False)
            outs.append(out)
        if not compact:
            g.defer_check()
        return torch.stack(outs, dim=1)

    @torch.no_grad()
    def predict(self, batch, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """Predictions of every model for every molecule of `batch` (PyG-layout `Data` / `Batch` on the device, or a
        compact store batch): `[M, K, P]`."""
        with M._device_guard(self.device):
            out = self._predict_all(batch)
        if getattr(batch, "is_compact", False):
            out = out[:batch.num_graphs]            # drop the padding graph (rows >= valid[1] belong to empty graphs)
        return out.to(dtype)

    @torch.no_grad()
    def predict_mean(self, batch, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """Ensemble mean per molecule, `[M, P]` - what `torch.concat(pred_para, dim=0).mean(0)` gives the reference
        for each molecule (evaluate_ensemble.py:146)."""
        return self.predict(batch).mean(dim=1).to(dtype)

    # ---- store-resident streams -------------------------------------------------------------------
    @torch.no_grad()
    def predict_from_store(self, store, ids: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """`[len(ids), K, P]` for graphs of a `CompactGraphStore`; the K forwards of a shape bucket are ONE captured
        CUDA graph."""
        from ..data.compact import bucket
        eng = self._engine
        with M._device_guard(self.device):
            n, s = store.sizes(ids)
            ent = eng._bucket("ensemble", store, int(ids.numel()), *bucket(n, s))
            store.batch(ids, out=ent["batch"])
            self._fresh_plans(ent["batch"])
            out = eng._run_bucket(ent, self._predict_all)
        return out[:int(ids.numel())].to(dtype)

    @torch.no_grad()
    def screen(self, store, lo: int = 0, hi: Optional[int] = None, max_nodes: int = 1 << 18,
               dtype: torch.dtype = torch.float32):
        """Ensemble mean `[hi - lo, P]` and all predictions `[hi - lo, K, P]` over a contiguous range of a store, in
        node-balanced micro-batches that share one shape bucket (one captured graph for the whole stream)."""
        from ..data.compact import MIN_PAD_NODES
        hi = store.num_graphs if hi is None else hi
        eng = self._engine
        with M._device_guard(self.device):
            total_nodes = int(store.node_count[lo:hi].sum())
            biggest = int(store.node_count[lo:hi].max())
            k = max(-(-total_nodes // max(int(max_nodes) - MIN_PAD_NODES - biggest, 1)), 1)
            budget = -(-total_nodes // k) + biggest
            pn = (budget + MIN_PAD_NODES + 127) // 128 * 128
            ps = (int(pn * float(store.slot_count[lo:hi].sum()) / max(total_nodes, 1) * 1.1) + 255) // 256 * 256
            ranges = store.node_balanced_ranges(budget, ps, lo, hi)
            cap = max(b - a for a, b in ranges)
            ent = eng._bucket("ensemble", store, cap, pn, ps)
            P = int(self.nets[0].mlp_params.num_para)
            out = torch.empty(hi - lo, len(self.nets), P, dtype=torch.float32, device=self.device)
            ids_all = torch.arange(lo, hi, dtype=torch.int64, device=self.device)
            for a, b in ranges:
                ids_dev = ids_all[a - lo:b - lo]
                store.batch(ids_dev, out=ent["batch"], ids_device=ids_dev)
                self._fresh_plans(ent["batch"])
                out[a - lo:b - lo].copy_(eng._run_bucket(ent, self._predict_all)[:b - a])
        return out.mean(dim=1).to(dtype), out.to(dtype)
