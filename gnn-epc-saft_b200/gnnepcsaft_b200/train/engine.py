"""Trainer-free training / screening loop for PNAPCSAFT on the CUDA kernels.

Reproduces the step the reference runs through Lightning (reference gnnepcsaft/train/models.py:
162-202 and SURVEY.md A.9): forward -> torchmetrics MAPE -> backward -> AdamW(amsgrad, eps 1e-5)
-> CosineAnnealingWarmRestarts(T_0 = warmup_steps) stepped every step, with

  * gradients in one flat fp32 arena laid out like the parameter arena, so data-parallel
    training issues exactly one NCCL all-reduce per step (the implicit Lightning DDP of
    reference train.py:132-156); the loss scalar rides in the same buffer (replaces the
    `sync_dist=True` scalar all-reduce of models.py:195-201);
  * the optimizer + LR schedule fused into one kernel over the arena;
  * BatchNorm statistics per rank (the reference has no SyncBatchNorm);
  * store-resident training (`step_from_store`, `fit_epoch`): the batch is gathered on the device by
    `gnn_compact_collate` into the static buffers of a shape bucket and the whole step (forward, loss, backward,
    optimizer; data parallel: up to the backward pass, the NCCL all-reduce and the optimizer kernel follow eagerly) is
    ONE captured CUDA graph per bucket - no per-step allocation, ~125 C-ABI calls and ~290 launches replaced by one
    `cudaGraphLaunch` (the reference's 512-graph batches, configs/default.py:22, are otherwise bound by host dispatch).
No autograd graph is built on this path.
"""

from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.distributed as dist

from .. import _lib
from . import models
from .models import _backward_impl, _c, _forward_impl, _require_cuda


class TrainEngine:
    def __init__(self, model, learning_rate: float = 1e-3, weight_decay: float = 1e-2, warmup_steps: int = 100,
                 betas=(0.9, 0.999), eps: float = 1e-5, process_group: Optional["dist.ProcessGroup"] = None,
                 distributed: Optional[bool] = None):
        self.net: models.PNAPCSAFT = model.model if isinstance(model, models.PNApcsaftL) else model
        self.lr, self.wd, self.t0 = float(learning_rate), float(weight_decay), int(warmup_steps)
        self.betas, self.eps = betas, float(eps)
        arena = self.net.arena()
        _require_cuda(arena, "the model")
        dev = arena.device
        self.numel = arena.numel()
        # +4: slot [numel] carries the step loss through the same all-reduce
        self.grad = torch.zeros(self.numel + 4, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(arena)
        self.exp_avg_sq = torch.zeros_like(arena)
        self.max_exp_avg_sq = torch.zeros_like(arena)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.group = process_group
        if distributed is None:
            distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(process_group) if distributed else 1
        if self.world > 1:     # identical initial parameters and buffers on every rank (DDP does the same)
            dist.broadcast(arena, src=0, group=process_group)
            for b in self.net.buffers():
                dist.broadcast(b, src=0, group=process_group)
        # captured CUDA graphs, one per (mode, graphs, padded nodes, padded slots) shape bucket; all share one memory
        # pool (they never run concurrently)
        self.use_cuda_graphs = os.environ.get("GNN_B200_CUDA_GRAPHS", "1") != "0" and dev.type == "cuda"
        self._buckets = {}
        self._pool = None
        self._warm = False
        self.graph_replays = 0
        self._arena_ptr = arena.data_ptr()
        self._copy_stream = None
        self._staging = {}

    @classmethod
    def from_config(cls, model, config, **kw):
        """`config` as in reference gnnepcsaft/configs/default.py (optimizer must be 'adam')."""
        if getattr(config, "optimizer", "adam") != "adam":
            raise ValueError("TrainEngine implements the reference's 'adam' (AdamW amsgrad) optimizer only")
        return cls(model, learning_rate=config.learning_rate, weight_decay=config.weight_decay,
                   warmup_steps=config.warmup_steps, **kw)

    @torch.no_grad()
    def step(self, batch) -> torch.Tensor:
        """One optimisation step on `batch` (needs `.para`); returns the (rank-mean) loss as a
        device scalar without synchronising."""
        with models._device_guard(self.grad.device):
            return self._step(batch)

    def _step(self, batch) -> torch.Tensor:
        loss = self._step_compute(batch)
        return self._step_update(loss)

    def _step_compute(self, batch) -> torch.Tensor:
        """Forward, loss and backward into the flat gradient arena (everything of a step that is rank-local)."""
        net = self.net
        net.arena()
        models.check_input_errors(block=False)
        compact = getattr(batch, "is_compact", False)
        valid = None
        if compact:     # store-resident compact batch (padded: graph `num_graphs` is the padding graph)
            out, sv = _forward_impl(net, None, None, None, None, batch.b, save=True, graph=batch, training=True)
            valid = batch.valid[1:2]
        else:
            batchvec = getattr(batch, "batch", None)
            num_graphs = batch.num_graphs if batchvec is not None else 1
            out, sv = _forward_impl(net, batch.x, batch.edge_index, batch.edge_attr, batchvec, int(num_graphs),
                                    save=True, training=True)
        target = batch.para
        if target is None:
            raise ValueError("TrainEngine.step: the batch has no `para` targets")
        _require_cuda(target, "batch.para")
        if target.dtype != torch.float32:
            target = target.float()
        target = target.contiguous()
        if target.numel() != out.numel():
            raise ValueError(f"batch.para has {target.numel()} values, expected {out.numel()}")
        dpred = torch.empty_like(out)
        _c("gnn_fill_zero", self.grad, self.grad.numel() * 4)
        loss = self.grad[self.numel:self.numel + 1]
        _c("gnn_mape_fwd_bwd", out, target, out.numel(), 1.0, loss, dpred, valid, int(out.shape[1]))
        _backward_impl(net, sv, dpred, self.grad)
        return loss

    def _step_update(self, loss: torch.Tensor) -> torch.Tensor:
        """Gradient all-reduce (one flat buffer, the loss rides along) + fused AdamW / schedule."""
        if self.world > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.group)
        scale = 1.0 / self.world
        _c("gnn_adamw_amsgrad_step", self.net._arena, self.grad, self.exp_avg, self.exp_avg_sq, self.max_exp_avg_sq,
           self.numel, self.step_count, self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.t0, scale)
        _c("gnn_step_increment", self.step_count)
        self.net._weights_epoch += 1        # invalidates the inference plan (weights changed through raw pointers)
        return loss * scale if self.world > 1 else loss

    def _train_bucket(self, ent):
        """One step on a bucket's static batch.  Single process: the whole step is one captured graph.  Data
        parallel: the captured graph ends with the backward pass and the collective + optimizer kernels (3 launches)
        are issued eagerly - NCCL registers the buffers of captured collectives across ranks at capture time, which
        needs lock-step captures, but the ranks meet new shape buckets at different steps."""
        self.net._weights_epoch += 1        # a replayed graph runs no Python: invalidate the inference plan here
        if self.world == 1:
            return self._run_bucket(ent, self._step)
        return self._step_update(self._run_bucket(ent, self._step_compute))

    @torch.no_grad()
    def predict(self, batch) -> torch.Tensor:
        """Eval-mode forward (running BatchNorm statistics), no autograd, no collective."""
        with models._device_guard(self.grad.device):      # the mode is passed down: no module-tree flips
            return self._predict(batch)

    def _fresh_plan(self, batch):
        """Captured eval graphs contain no weight-preparation kernels: bring the model's inference plan up to date with
        one eager forward whenever the weights changed since it was filled."""
        net = self.net
        kind = "tiles" if (models.TILE_INFER and getattr(batch, "tileable", False)) else "layers"
        if net._matmul_mode() >= 0 and net._infer_plan(self.grad.device, kind).filling:
            self._predict(batch)

    def _predict(self, batch) -> torch.Tensor:
        net = self.net
        if getattr(batch, "is_compact", False):
            out, _ = _forward_impl(net, None, None, None, None, batch.b, save=False, graph=batch, training=False)
            return out          # [num_graphs + 1, P]: the last row is the padding graph
        batchvec = getattr(batch, "batch", None)
        num_graphs = batch.num_graphs if batchvec is not None else 1
        out, _ = _forward_impl(net, batch.x, batch.edge_index, batch.edge_attr, batchvec, int(num_graphs), save=False,
                               training=False)
        return out

    # ---- store-resident path: compact device collate + one captured CUDA graph per shape bucket --------------
    def _bucket(self, mode: str, store, nb: int, pn: int, ps: int):
        key = (mode, store.uid, nb, pn, ps)
        ent = self._buckets.get(key)
        if ent is None:
            ent = {"batch": store.allocate(nb, pn, ps), "graph": None, "out": None}
            self._buckets[key] = ent
        return ent

    def _run_bucket(self, ent, fn):
        """Run `fn(static batch)` for a bucket: eagerly until the process is warm (lazy constants, workspaces and
        kernel attributes exist), then captured once and replayed."""
        if not self.use_cuda_graphs:
            return fn(ent["batch"])
        arena_ptr = self.net.arena().data_ptr()
        if arena_ptr != self._arena_ptr:
            # the parameter arena was rebuilt (model.to(...), .float(), ...): every captured graph still addresses
            # the old one - drop them all (the static batch buffers stay valid)
            for other in self._buckets.values():
                other["graph"], other["out"] = None, None
            self._arena_ptr = arena_ptr
        if ent["graph"] is None:
            if not self._warm:
                self._warm = True
                return fn(ent["batch"])
            torch.cuda.synchronize(self.grad.device)
            g = torch.cuda.CUDAGraph()
            if self._pool is None:
                self._pool = torch.cuda.graph_pool_handle()
            # thread_local: other threads (NCCL watchdog, copy threads) may keep issuing CUDA calls while capturing
            with torch.cuda.graph(g, pool=self._pool, capture_error_mode="thread_local"):
                ent["out"] = fn(ent["batch"])
            ent["graph"] = g
        ent["graph"].replay()
        self.graph_replays += 1
        return ent["out"]

    @torch.no_grad()
    def step_from_store(self, store, ids: torch.Tensor, ids_device: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One optimisation step on the graphs `ids` (host int64 tensor) of a `CompactGraphStore`.  Returns the loss
        as a device scalar that the NEXT step of the same bucket overwrites (read or clone it before)."""
        from ..data.compact import bucket
        with models._device_guard(self.grad.device):
            self.net.arena()
            n, s = store.sizes(ids)
            pn, ps = bucket(n, s)
            ent = self._bucket("train", store, int(ids.numel()), pn, ps)
            store.batch(ids, out=ent["batch"], ids_device=ids_device)
            return self._train_bucket(ent)

    @torch.no_grad()
    def step_from_host(self, store, host_batch) -> torch.Tensor:
        """The same for a compact batch assembled on the host (`store.host_batch(ids)`, pinned): one H2D copy per
        field into the bucket's static buffers, then the captured step."""
        with models._device_guard(self.grad.device):
            self.net.arena()
            ent = self._bucket("train", store, host_batch.num_graphs, host_batch.pad_nodes, host_batch.pad_slots)
            ent["batch"].copy_from(host_batch)
            return self._train_bucket(ent)

    @torch.no_grad()
    def steps_from_host(self, store, host_batches, depth: int = 2):
        """Generator: one optimisation step per pinned host compact batch (`store.host_batch(ids)`), yielding each
        step's loss.  The H2D copies of the next `depth` batches run on a copy stream into PREALLOCATED staging buffers
        (per shape bucket, reused round-robin behind events) while the current step computes; the step then copies its
        batch device-to-device into the bucket's static buffers and replays the captured graph.  No per-step
        allocation: with tensors allocated per batch on the copy stream (DevicePrefetcher) the caching allocator's
        deferred frees degraded into cudaMalloc calls after a few dozen steps (4.8 -> 5.1-7.6 ms per step)."""
        from collections import deque
        dev = self.grad.device
        with models._device_guard(dev):
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
            copy_stream = self._copy_stream
            source = iter(host_batches)
            queue = deque()

            def enqueue() -> bool:
                hb = next(source, None)
                if hb is None:
                    return False
                key = (hb.num_graphs, hb.pad_nodes, hb.pad_slots)
                ring = self._staging.get((store.uid, key))
                if ring is None:
                    ring = {"bufs": [store.allocate(*key) for _ in range(depth + 1)], "free": [None] * (depth + 1), "next": 0}
                    self._staging[(store.uid, key)] = ring
                j = ring["next"]
                ring["next"] = (j + 1) % len(ring["bufs"])
                with torch.cuda.stream(copy_stream):
                    if ring["free"][j] is not None:
                        copy_stream.wait_event(ring["free"][j])     # its previous contents were consumed
                    ring["bufs"][j].copy_from(hb)
                    ready = torch.cuda.Event()
                    ready.record(copy_stream)
                queue.append((ring, j, key, ready))
                return True

            while len(queue) < depth and enqueue():
                pass
            while queue:
                ring, j, key, ready = queue.popleft()
                cur = torch.cuda.current_stream(dev)
                cur.wait_event(ready)
                self.net.arena()
                ent = self._bucket("train", store, *key)
                ent["batch"].copy_from(ring["bufs"][j])
                free = torch.cuda.Event()
                free.record(cur)
                ring["free"][j] = free
                enqueue()
                yield self._train_bucket(ent)

    @torch.no_grad()
    def predict_from_store(self, store, ids: torch.Tensor, ids_device: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Eval-mode forward of the graphs `ids` of a `CompactGraphStore`: [len(ids), P] (a view of the bucket's
        static output, overwritten by the next call on the same bucket)."""
        from ..data.compact import bucket
        net = self.net
        with models._device_guard(self.grad.device):
            net.arena()
            n, s = store.sizes(ids)
            pn, ps = bucket(n, s)
            ent = self._bucket("eval", store, int(ids.numel()), pn, ps)
            store.batch(ids, out=ent["batch"], ids_device=ids_device)
            self._fresh_plan(ent["batch"])
            out = self._run_bucket(ent, self._predict)
        return out[:int(ids.numel())]

    @torch.no_grad()
    def screen(self, store, lo: int = 0, hi: Optional[int] = None, max_nodes: int = 1 << 19, out=None,
               passes: int = 1):
        """Batched inference over the graph range [lo, hi) of a store (the reference's per-molecule prediction loops,
        evaluate_ensemble.py:128-208 / demo/utils.py:20-40, as one stream): micro-batches of consecutive graphs with
        at most `max_nodes` nodes, all padded to ONE bucket, so the whole stream replays one captured CUDA graph.
        `out=False` discards the predictions (throughput runs).  Returns (predictions [hi - lo, P] | None, number of
        micro-batches, node counts of the micro-batches)."""
        from ..data.compact import MIN_PAD_NODES
        hi = store.num_graphs if hi is None else hi
        net = self.net
        dev = self.grad.device
        if True:
            with models._device_guard(dev):
                net.arena()
                # even micro-batches: k = ceil(nodes / budget) ranges of about nodes / k each, so that the last one is
                # not mostly padding; the slot capacity is generous (unowned surplus slots cost nothing)
                total_nodes = int(store.node_count[lo:hi].sum())
                biggest = int(store.node_count[lo:hi].max())
                k = max(-(-total_nodes // max(int(max_nodes) - MIN_PAD_NODES - biggest, 1)), 1)
                budget = -(-total_nodes // k) + biggest
                pn = (budget + MIN_PAD_NODES + 127) // 128 * 128
                avg_slots = float(store.slot_count[lo:hi].sum()) / max(total_nodes, 1)
                ps = (int(pn * avg_slots * 1.1) + 255) // 256 * 256
                ranges = store.node_balanced_ranges(budget, ps, lo, hi)
                cap = max(b - a for a, b in ranges)
                if out is None:
                    out = torch.empty(hi - lo, int(net.mlp_params.num_para), dtype=torch.float32, device=dev)
                ent = self._bucket("screen", store, cap, pn, ps)
                ids_all = torch.arange(lo, hi, dtype=torch.int64, device=dev)
                for _ in range(passes):
                    for a, b in ranges:
                        ids_dev = ids_all[a - lo:b - lo]
                        store.batch(ids_dev, out=ent["batch"], ids_device=ids_dev)
                        self._fresh_plan(ent["batch"])
                        y = self._run_bucket(ent, self._predict)
                        if out is not False:
                            out[a - lo:b - lo].copy_(y[:b - a])
        nodes = [int(store.node_count[a:b].sum()) for a, b in ranges]
        return (None if out is False else out), len(ranges), nodes

    def fit_epoch(self, store, batch_size: int, shuffle: bool = True, generator=None, drop_last: bool = False,
                  rank: int = 0, world: Optional[int] = None):
        """One epoch over the store (the reference's `trainer.fit` inner loop, train.py:186-201): returns the mean of
        the per-step losses as a python float (one synchronisation per epoch)."""
        world = self.world if world is None else world
        total = torch.zeros(1, dtype=torch.float32, device=self.grad.device)
        steps = 0
        for ids in store.epoch(batch_size, shuffle=shuffle, generator=generator, drop_last=drop_last, rank=rank,
                               world=world):
            total += self.step_from_store(store, ids)
            steps += 1
        return float(total) / max(steps, 1)

    def optimizer_state_dict(self):
        """torch.optim.AdamW-style state for the legacy `.pth` checkpoints (reference train/utils.py:109-119)."""
        state = {}
        for i, (p, off) in enumerate(zip(self.net._param_list(), self.net._arena_offsets)):
            sl = slice(off, off + p.numel())
            state[i] = {"step": self.step_count.clone().float().reshape(()),
                        "exp_avg": self.exp_avg[sl].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[sl].view(p.shape).clone(),
                        "max_exp_avg_sq": self.max_exp_avg_sq[sl].view(p.shape).clone()}
        t = int(self.step_count.item())
        cur_lr = 0.5 * self.lr * (1.0 + math.cos(math.pi * (t % self.t0) / self.t0))     # after t scheduler steps
        group = {"lr": cur_lr, "initial_lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.wd, "amsgrad": True,
                 "params": list(range(len(state)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, state_dict):
        """Inverse of `optimizer_state_dict`: restores the Adam moments, the amsgrad maxima and the step counter
        (which also fixes the phase of the cosine-restart schedule), so a resumed run continues where it stopped
        (reference train/utils.py:220-233, train.py:158-184)."""
        state = state_dict["state"]
        params = self.net._param_list()
        if len(state) != len(params):
            raise ValueError(f"optimizer state has {len(state)} entries, the model has {len(params)} parameters")
        step = 0
        with torch.no_grad():
            for i, (p, off) in enumerate(zip(params, self.net._arena_offsets)):
                st = state[i] if i in state else state[str(i)]
                sl = slice(off, off + p.numel())
                self.exp_avg[sl].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[sl].copy_(st["exp_avg_sq"].reshape(-1))
                self.max_exp_avg_sq[sl].copy_(st["max_exp_avg_sq"].reshape(-1))
                step = int(st["step"])
            self.step_count.fill_(step)
        groups = state_dict.get("param_groups") or [{}]
        self.wd = float(groups[0].get("weight_decay", self.wd))
        self.lr = float(groups[0].get("initial_lr", self.lr))

    def scheduler_state_dict(self):
        """`CosineAnnealingWarmRestarts(T_0 = warmup_steps)` state after `step_count` per-step calls."""
        t = int(self.step_count.item())
        return {"T_0": self.t0, "T_i": self.t0, "T_mult": 1, "eta_min": 0, "T_cur": t % self.t0, "last_epoch": t,
                "base_lrs": [self.lr]}

    def save_checkpoint(self, path: str, step=None):
        """`.pth` (reference train/utils.py:109-119 layout) or Lightning-style `.ckpt`, chosen by the extension."""
        from . import utils
        step = int(self.step_count.item()) if step is None else step
        if str(path).endswith(".ckpt"):
            utils.save_lightning_checkpoint(self.net, path, engine=self, global_step=step)
        else:
            utils.savemodel(self.net, self, None, path, step)

    def load_checkpoint(self, path: str, load_optimizer: bool = True):
        """Reads either layout back (model weights, and optimizer / schedule state unless `load_optimizer=False`,
        the reference's `config.change_opt`, train.py:170-175)."""
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        if "model_state_dict" in ckpt:
            self.net.load_state_dict(ckpt["model_state_dict"])
            opt = ckpt.get("optimizer_state_dict")
        else:
            sd = {k[len("model."):] if k.startswith("model.") else k: v for k, v in ckpt["state_dict"].items()}
            self.net.load_state_dict(sd)
            opt = (ckpt.get("optimizer_states") or [None])[0]
        if load_optimizer and opt is not None:
            self.load_optimizer_state_dict(opt)
        return ckpt.get("step", ckpt.get("global_step"))
