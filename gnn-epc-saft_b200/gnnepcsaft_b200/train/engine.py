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

import math
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
        with models._device_guard(self.grad.device):
            return self._step(batch)

    def _step(self, batch) -> torch.Tensor:
        net = self.net
        net.train()
        arena = net.arena()
        models.check_input_errors(block=False)
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
          with models._device_guard(self.grad.device):
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
