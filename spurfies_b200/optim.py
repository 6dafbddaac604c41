"""the optimiser step that follows the hot path, as two HBM-bound kernels over flat buffers.

Reference (spurfies/train.py): ``torch.optim.Adam(trainable, lr=5e-4)`` (:168-189), ``CosineAnnealingLR(T_max=100_000,
eta_min=3e-4)`` (:191-193), and per step ``zero_grad -> backward -> clip_grad_norm_(1.0) -> on_after_backward (NaN/Inf
guard, :548-564) -> optimizer.step -> scheduler.step`` (:355-363).

``FusedAdam`` keeps every trainable tensor as a view of one flat fp32 parameter buffer (``p.data`` is re-pointed, the
module's state_dict is unchanged), every ``p.grad`` as a view of one flat gradient buffer (so autograd accumulates in
place and the data-parallel all-reduce needs no packing), and the two Adam moments flat as well.  ``step()`` is
``spf_grad_sumsq`` + ``spf_adam_step``: global norm, clip coefficient, NaN/Inf guard, Adam update and the gradient
clear in one read of (p, g, m, v) and one write of (p, m, v, g).  No host synchronisation: the step count and the
learning rate live in a 2-float device buffer, so the whole training step stays CUDA-graph capturable.

``state_dict`` / ``load_state_dict`` use torch.optim.Adam's layout (per-parameter ``step`` / ``exp_avg`` /
``exp_avg_sq``), so the reference's ``OptimizerParameters/*.pth`` checkpoints (train.py:300-328) load unchanged.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional

import torch

from . import _lib
from ._lib import call, ptr, stream
from .dist import flat_offsets


def cosine_lr(step: int, base_lr: float = 5.0e-4, eta_min: float = 3.0e-4, t_max: int = 100_000) -> float:
    """Closed form of torch.optim.lr_scheduler.CosineAnnealingLR after `step` scheduler steps (train.py:191-193)."""
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * step / t_max)) / 2.0


class FusedAdam:
    _LR_SLOTS = 64

    def __init__(self, params: Iterable[torch.Tensor], lr: float = 5.0e-4, betas=(0.9, 0.999), eps: float = 1.0e-8,
                 max_norm: float = 1.0, grad_flat: Optional[torch.Tensor] = None):
        self.params: List[torch.Tensor] = [p for p in params]
        if not self.params:
            raise ValueError("FusedAdam: no parameters")
        p0 = self.params[0]
        if not p0.is_cuda:
            raise _lib.SpfError("FusedAdam needs CUDA parameters (there is no CPU fallback)")
        dev = p0.device
        self.lr, self.betas, self.eps, self.max_norm = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(max_norm)
        # 16-byte aligned segments so the kernels can use 128-bit accesses on every tensor boundary
        assert all(p.dtype == torch.float32 for p in self.params)
        self.offsets, off = flat_offsets(self.params, align=4)
        self.numel = off
        self.flat_p = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_g = grad_flat if grad_flat is not None else torch.zeros(off, dtype=torch.float32, device=dev)
        assert self.flat_g.numel() == off and self.flat_g.is_cuda
        self.exp_avg = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=dev)
        self.state = torch.tensor([0.0, self.lr], dtype=torch.float32, device=dev)   # [steps taken, lr]
        # ring of pinned staging slots for the asynchronous lr upload: the host may run many (graph-replayed) steps
        # ahead of the GPU, so a single pinned word could be overwritten before its copy has been executed
        self._lr_host = torch.full((self._LR_SLOTS,), self.lr, dtype=torch.float32).pin_memory()
        self._lr_events = [None] * self._LR_SLOTS
        self._lr_n = 0
        self.norm_sq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.info = torch.zeros(2, dtype=torch.float32, device=dev)
        self._ws = torch.empty(_lib.lib.spf_optim_workspace_bytes(), dtype=torch.uint8, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.flat_p[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
        self.attach_grads()

    # ------------------------------------------------------------------ views
    def view_of(self, flat: torch.Tensor, i: int) -> torch.Tensor:
        p, o = self.params[i], self.offsets[i]
        return flat[o:o + p.numel()].view_as(p)

    def attach_grads(self) -> None:
        for i, p in enumerate(self.params):
            p.grad = self.view_of(self.flat_g, i)

    def grads_attached(self) -> bool:
        base = self.flat_g.data_ptr()
        return all(p.grad is not None and p.grad.data_ptr() == base + 4 * o and p.grad.is_contiguous()
                   for p, o in zip(self.params, self.offsets))

    def zero_grad(self) -> None:
        self.flat_g.zero_()

    def set_lr(self, lr: float) -> None:
        """Learning rate of the next step (host scheduler -> one 4-byte async H2D copy, outside any captured graph)."""
        self.lr = float(lr)
        i = self._lr_n % self._LR_SLOTS
        self._lr_n += 1
        if self._lr_events[i] is not None:
            self._lr_events[i].synchronize()   # the copy that last used this slot has run (64 steps ago: never waits)
        self._lr_host[i] = self.lr
        self.state[1:2].copy_(self._lr_host[i:i + 1], non_blocking=True)
        if self._lr_events[i] is None:
            self._lr_events[i] = torch.cuda.Event()
        self._lr_events[i].record()

    # ------------------------------------------------------------------ step
    def step(self, grad_scale: float = 1.0, zero_grad: bool = True) -> None:
        """clip_grad_norm_(max_norm) + NaN/Inf guard + Adam + (optionally) zero_grad on `grad_scale * grad`."""
        call("spf_grad_sumsq", ptr(self.flat_g), self.numel, float(grad_scale), ptr(self.norm_sq), ptr(self._ws),
             self._ws.numel(), stream())
        call("spf_adam_step", ptr(self.flat_p), ptr(self.flat_g), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.numel,
             ptr(self.norm_sq), ptr(self.state), float(grad_scale), float(self.max_norm), self.betas[0], self.betas[1],
             float(self.eps), int(zero_grad), ptr(self.info), stream())

    def total_norm(self) -> torch.Tensor:
        return self.info[0]

    def skipped(self) -> torch.Tensor:
        return self.info[1]

    # ------------------------------------------------------------------ torch.optim.Adam-compatible checkpoints
    def state_dict(self) -> Dict:
        steps = self.state[0].detach().clone()
        st = {i: {"step": steps.clone(), "exp_avg": self.view_of(self.exp_avg, i).clone(),
                  "exp_avg_sq": self.view_of(self.exp_avg_sq, i).clone()} for i in range(len(self.params))}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(self.params)))}
        return {"state": st, "param_groups": [group]}

    def load_state_dict(self, sd: Dict) -> None:
        ids = [i for g in sd["param_groups"] for i in g["params"]]
        if len(ids) != len(self.params):
            raise ValueError(f"optimizer state has {len(ids)} parameters, expected {len(self.params)}")
        step = 0.0
        with torch.no_grad():
            for i, pid in enumerate(ids):
                s = sd["state"].get(pid)
                if s is None:
                    continue
                self.view_of(self.exp_avg, i).copy_(s["exp_avg"])
                self.view_of(self.exp_avg_sq, i).copy_(s["exp_avg_sq"])
                step = max(step, float(s["step"]))
            groups = [g for g in sd["param_groups"] if g["params"]]
            if groups:
                self.betas = tuple(float(b) for b in groups[0]["betas"])
                self.eps = float(groups[0]["eps"])
                self.lr = float(groups[0]["lr"])
            self.state.copy_(torch.tensor([step, self.lr], dtype=torch.float32))
