"""Trainer-free training / screening loop for PNAPCSAFT on the CUDA kernels.

Reproduces the step the reference runs through Lightning (reference gnnepcsaft/train/models.py:
162-202 and SURVEY.md A.9): forward -> torchmetrics MAPE -> backward -> AdamW(amsgrad, eps 1e-5)
-> CosineAnnealingWarmRestarts(T_0 = warmup_steps) stepped every step, with

  * gradients in one flat fp32 arena laid out like the parameter arena, so data-parallel
    training issues exactly one NCCL all-reduce per step (the implicit Lightning DDP of
    reference train.py:132-156); the loss scalar rides in the same buffer (replaces the
    `sync_dist=True` scalar all-reduce of models.py:195-201);
  * the optimizer + LR schedule fused into one kernel over the arena;
  * BatchNorm statistics per rank (the reference has no SyncBatchNorm).
No autograd graph is built on this path.
"""

from __future__ import annotations

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
        net = self.net
        net.train()
        arena = net.arena()
        batchvec = getattr(batch, "batch", None)
        num_graphs = batch.num_graphs if batchvec is not None else 1
        out, sv = _forward_impl(net, batch.x, batch.edge_index, batch.edge_attr, batchvec, int(num_graphs), save=True)
        target = batch.para
        _require_cuda(target, "batch.para")
        if target.dtype != torch.float32:
            target = target.float()
        target = target.contiguous()
        if target.numel() != out.numel():
            raise ValueError(f"batch.para has {target.numel()} values, expected {out.numel()}")
        dpred = torch.empty_like(out)
        _c("gnn_fill_zero", self.grad, self.grad.numel() * 4)
        loss = self.grad[self.numel:self.numel + 1]
        _c("gnn_mape_fwd_bwd", out, target, out.numel(), 1.0, loss, dpred)
        _backward_impl(net, sv, dpred, self.grad)
        if self.world > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.group)
        scale = 1.0 / self.world
        _c("gnn_adamw_amsgrad_step", arena, self.grad, self.exp_avg, self.exp_avg_sq, self.max_exp_avg_sq, self.numel,
           self.step_count, self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.t0, scale)
        _c("gnn_step_increment", self.step_count)
        return loss * scale if self.world > 1 else loss

    @torch.no_grad()
    def predict(self, batch) -> torch.Tensor:
        """Eval-mode forward (running BatchNorm statistics), no autograd, no collective."""
        net = self.net
        was_training = net.training
        net.eval()
        try:
            batchvec = getattr(batch, "batch", None)
            num_graphs = batch.num_graphs if batchvec is not None else 1
            out, _ = _forward_impl(net, batch.x, batch.edge_index, batch.edge_attr, batchvec, int(num_graphs),
                                   save=False)
        finally:
            net.train(was_training)
        return out

    def optimizer_state_dict(self):
        """torch.optim.AdamW-style state for the legacy `.pth` checkpoints (reference train/utils.py:109-119)."""
        state = {}
        for i, (p, off) in enumerate(zip(self.net._param_list(), self.net._arena_offsets)):
            sl = slice(off, off + p.numel())
            state[i] = {"step": self.step_count.clone().float().reshape(()),
                        "exp_avg": self.exp_avg[sl].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[sl].view(p.shape).clone(),
                        "max_exp_avg_sq": self.max_exp_avg_sq[sl].view(p.shape).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.wd, "amsgrad": True,
                 "params": list(range(len(state)))}
        return {"state": state, "param_groups": [group]}
