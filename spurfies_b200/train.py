"""One optimisation step of the reference trainer around the hot path (spurfies/train.py:330-397): forward,
VolSDFLoss, backward, clip_grad_norm_(1.0), NaN guard, Adam (lr 5e-4 on every trainable tensor, train.py:168-189),
cosine learning-rate schedule (train.py:191-193).  Clip + guard + Adam + zero_grad are the two kernels of
`spurfies_b200/optim.py::FusedAdam` over flat parameter / gradient / moment buffers.

Data parallel (new functionality): one process per GPU, the step's rays are sharded across
ranks, the neural points / grid / latents / MLPs are replicated, and ONE NCCL all-reduce per step carries the
flat fp32 gradient buffer (latents N x 96 + colour MLP + radiance head + beta).  Clipping uses the global norm,
so it runs after the all-reduce; every rank then takes the identical Adam step.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import fields
from .dist import FlatGradReducer
from .model import PointVolSDF, VolSDFLoss
from .optim import FusedAdam, cosine_lr


def _clone_tree(v):
    """Static copy of a (nested) input dict: tensors are cloned, dicts recursed, everything else kept."""
    if torch.is_tensor(v):
        return v.clone()
    if isinstance(v, dict):
        return {k: _clone_tree(x) for k, x in v.items()}
    return v


def _like_tree(v):
    if torch.is_tensor(v):
        return torch.empty_like(v)
    if isinstance(v, dict):
        return {k: _like_tree(x) for k, x in v.items()}
    return v


def _copy_tree(dst, src, path="") -> None:
    """dst[k].copy_(src[k]) for every tensor of a (nested) dict.  The structure and shapes must be the captured ones:
    a CUDA graph bakes in the device pointers, so an entry that cannot be copied into its static buffer is an error,
    never silently skipped."""
    for k, v in src.items():
        if torch.is_tensor(v):
            if k not in dst or not torch.is_tensor(dst[k]) or dst[k].shape != v.shape:
                raise ValueError(f"graph replay: input '{path}{k}' does not match the captured step "
                                 f"({tuple(v.shape)} vs {tuple(dst[k].shape) if k in dst and torch.is_tensor(dst[k]) else None})")
            dst[k].copy_(v, non_blocking=True)
        elif isinstance(v, dict):
            if not isinstance(dst.get(k), dict):
                raise ValueError(f"graph replay: input '{path}{k}' was not a dict when the step was captured")
            _copy_tree(dst[k], v, path + k + ".")
        elif v is None and dst.get(k) is not None:
            raise ValueError(f"graph replay: input '{path}{k}' is None but the captured step had a value")


def _to_device_tree(v, dev):
    if torch.is_tensor(v):
        return v.to(dev, non_blocking=True)
    if isinstance(v, dict):
        return {k: _to_device_tree(x, dev) for k, x in v.items()}
    return v


def _record_tree(v, stream) -> None:
    if torch.is_tensor(v):
        if v.is_cuda:
            v.record_stream(stream)
    elif isinstance(v, dict):
        for x in v.values():
            _record_tree(x, stream)


class TrainStep:
    def __init__(self, model: PointVolSDF, lr: float = 5.0e-4, grad_clip: float = 1.0, loss: Optional[VolSDFLoss] = None,
                 world_size: int = 1, lr_schedule: bool = True, grad_compress: Optional[str] = None, dp_overlap: bool = False):
        """grad_compress="bf16" (opt-in, bf16 precision mode only): the latent tables' gradients -- 96 % of the exchanged
        bytes -- are all-reduced as bf16 (FlatGradReducer.bf16_prefix); default None = exact fp32 exchange.
        dp_overlap (opt-in; world_size > 1, fp32 exchange): the colour latents' gradient (64 % of the exchanged bytes) is
        all-reduced on NCCL's stream as soon as the colour field's backward is done, under the geometry backward and the
        regulariser that follow; the rest is reduced after the backward.  Same sums, same step (tests/test_gpu_dist.py).
        Off by default because it does not pay on NVSwitch: measured on 8 B200 (profiles/r03c_*): no exchange at all
        4.12 ms/step, one all-reduce after the backward 4.30 ms (40 MB through NVLS at NCCL's own 411 GB/s model = 0.12 ms
        + latency), early reduction started before the weight-gradient kernel 4.57 ms (NCCL's copy kernels next to a kernel
        that runs at 90 % of the HBM peak), started after it: no difference beyond run-to-run noise at 2 GPUs."""
        self.model = model
        self.loss = loss or VolSDFLoss()
        for prm in list(model.F_geometry.parameters()) + list(model.T.parameters()):
            prm.requires_grad_(False)  # the local-prior SDF field is frozen (train.py:151-154)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.grad_clip = grad_clip
        self.world_size = world_size
        model.set_data_parallel(world_size)   # count-normalised loss terms use global counts (PointVolSDF.forward)
        self.base_lr, self.lr_schedule, self.iter_step = lr, lr_schedule, 0
        n_half = 0
        if grad_compress is not None:
            if grad_compress != "bf16" or model.precision != "bf16":
                raise ValueError("grad_compress: only 'bf16', and only with precision='bf16'")
            lat = [p for p in self.params if p is model.neural_feats_color or p is model.neural_feats_geometry]
            assert all(a is b for a, b in zip(lat, self.params)), "the latent tables must be the first parameters"
            n_half = sum((p.numel() + 3) // 4 * 4 for p in lat)
        self._reducer = FlatGradReducer(self.params, world_size, align=4, bf16_prefix=n_half)
        # the colour field's split-K weight-gradient launch runs on a side stream under the rest of the backward
        # (fields.WGRAD_SIDE); SPF_WGRAD_SIDE=0 keeps it in line
        self.wgrad_side = os.environ.get("SPF_WGRAD_SIDE", "1") != "0"
        self._early_n, self._early_work = 0, None
        if world_size > 1 and dp_overlap and grad_compress is None and self.params[0] is model.neural_feats_color:
            self._early_n = self._reducer.offsets[1] if len(self.params) > 1 else self._reducer.numel
        # diagnosis only (bench.py marks the line): leave the gradient exchange out to time the step without it
        self._diag_skip_reduce = world_size > 1 and os.environ.get("SPF_DP_DIAG_SKIP_REDUCE") == "1"
        # every p.data / p.grad becomes a view of a flat buffer; the gradient buffer is the one NCCL reduces
        self.opt = FusedAdam(self.params, lr=lr, max_norm=grad_clip if grad_clip else 0.0, grad_flat=self._reducer.flat())
        self._graph = None
        self._static = None
        self.graph_error = None

    def _early_reduce(self):
        """Fired by fields.ColorField.backward: the first `_early_n` elements of the flat gradient (the colour latents)
        are final.  Asynchronous all-reduce: NCCL's stream waits for the kernels enqueued so far, this stream goes on."""
        if self._early_work is not None:
            raise RuntimeError("the colour field ran two backwards in one step: its gradient was already being reduced")
        self._early_work = dist.all_reduce(self._reducer.flat()[:self._early_n], op=dist.ReduceOp.SUM,
                                           group=self._reducer.group, async_op=True)

    def _allreduce_grads(self):
        if self._diag_skip_reduce:
            return
        if self._early_work is None:
            self._reducer.reduce(average=False)   # 1/world is folded into the optimiser's clip coefficient
            return
        flat = self._reducer.flat()
        if self._early_n < flat.numel():
            dist.all_reduce(flat[self._early_n:], op=dist.ReduceOp.SUM, group=self._reducer.group)
        self._early_work.wait()   # this stream waits for the early part (no host synchronisation)
        self._early_work = None

    def _tick_lr(self):
        """scheduler.step() (train.py:363): host-side closed form, one 4-byte H2D copy; outside the captured graph."""
        if self.lr_schedule:
            self.opt.set_lr(cosine_lr(self.iter_step, self.base_lr))
        self.iter_step += 1

    # ------------------------------------------------------------------ CUDA-graph replay of the whole step
    def capture(self, batch, gt, rng) -> bool:
        """Capture forward + loss + backward + (all-reduce) + clip + Adam into one CUDA graph.  Possible in both
        precision modes because the step has no host synchronisation (slot counts stay on the device) and every big
        buffer is persistent.
        Inputs are copied into static tensors before each replay.  Returns False (and stays eager) if capture fails."""
        try:
            import gc
            self.model._last = None
            gc.collect()  # drop every reference to earlier eager steps' autograd graphs (see model.forward)
            # every tensor of the inputs -- including the nested ``local_data`` dict (feature maps and cameras of the
            # step's view, which change every step) -- gets a static device copy the graph reads from
            dev = self.opt.flat_p.device
            self._static = tuple(_clone_tree(_to_device_tree(d, dev)) for d in (batch, gt, rng))
            # The warm-up steps below are real optimisation steps on one batch: snapshot everything they change
            # (parameters, both Adam moments, the step count / lr, the gradient buffer) and restore it afterwards, so
            # that capture() has no effect on the training trajectory.
            opt = self.opt
            snap = [t.clone() for t in (opt.flat_p, opt.exp_avg, opt.exp_avg_sq, opt.state, opt.flat_g)]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._eager(*self._static)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            restore = lambda: [t.copy_(s_) for t, s_ in zip((opt.flat_p, opt.exp_avg, opt.exp_avg_sq, opt.state, opt.flat_g), snap)]
            restore()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._static_out = self._eager(*self._static)
            restore()   # (capture does not execute, but keep the invariant explicit and cheap)
            torch.cuda.synchronize()
            self._graph = g
            return True
        except Exception as e:  # noqa: BLE001 - report and fall back to eager
            import os, traceback
            if os.environ.get("SPF_DEBUG_GRAPH"):
                traceback.print_exc()
            self._graph = None
            self.graph_error = f"{type(e).__name__}: {e}"
            torch.cuda.synchronize()
            return False

    def replay(self, batch, gt, rng):
        for st, src in zip(self._static, (batch, gt, rng)):
            _copy_tree(st, src)
        if getattr(self, "_copy_stream", None) is not None:
            self._free.record()      # the staging set (if that is where the inputs came from) may be refilled
        self._tick_lr()
        self._graph.replay()
        return self._static_out

    # ------------------------------------------------------------------ input prefetch (overlaps H2D with the previous step)
    def prefetch(self, batch, gt, rng) -> None:
        """Start copying the NEXT step's (pinned host) inputs to the device on a side stream while the current step runs;
        ``step_prefetched()`` then consumes them.  With a captured graph the copies land in a staging set that the step
        moves into the graph's static inputs with device-to-device copies (3 MB, microseconds)."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
            self._ready, self._free = torch.cuda.Event(), torch.cuda.Event()
            self._free.record()
            self._stage = None
        cs = self._copy_stream
        if self._graph is None:
            with torch.cuda.stream(cs):
                dev = self.opt.flat_p.device
                self._pending = tuple(_to_device_tree(d, dev) for d in (batch, gt, rng))
                self._ready.record(cs)
            return
        if self._stage is None:
            self._stage = tuple(_like_tree(d) for d in self._static)
        cs.wait_event(self._free)   # the previous step has moved the staging set into the static inputs
        with torch.cuda.stream(cs):
            for st, src in zip(self._stage, (batch, gt, rng)):
                _copy_tree(st, src)
            self._ready.record(cs)
        self._pending = self._stage

    def step_prefetched(self) -> Dict[str, torch.Tensor]:
        """Run one step on the inputs given to the last ``prefetch`` call."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._ready)
        b, g, r = self._pending
        if self._graph is None:
            for d in (b, g, r):
                _record_tree(d, cur)
            self._tick_lr()
            return self._eager(b, g, r)
        out = self.replay(b, g, r)   # device-to-device copies into the static inputs, then the graph
        return out

    # ------------------------------------------------------------------ checkpoints (train.py:300-328 saves the scheduler too)
    def state_dict(self) -> Dict:
        return {"iter_step": int(self.iter_step), "base_lr": float(self.base_lr), "optimizer": self.opt.state_dict()}

    def load_state_dict(self, sd: Dict) -> None:
        self.opt.load_state_dict(sd["optimizer"])
        self.base_lr = float(sd.get("base_lr", self.base_lr))
        # without a saved scheduler position, resume where the optimiser's step count says (one tick per step)
        self.iter_step = int(sd["iter_step"]) if "iter_step" in sd else int(self.opt.state[0].item())
        if self.lr_schedule:
            self.opt.set_lr(cosine_lr(max(self.iter_step - 1, 0), self.base_lr))

    def __call__(self, batch, gt, rng=None) -> Dict[str, torch.Tensor]:
        if self._graph is not None:
            return self.replay(batch, gt, rng)
        self._tick_lr()
        return self._eager(batch, gt, rng)

    def _eager(self, batch: Dict[str, torch.Tensor], gt: Dict[str, torch.Tensor], rng=None) -> Dict[str, torch.Tensor]:
        self.model.train()
        out = self.model(batch, fast=1, rng=rng, dense_outputs=True)  # fast=1: train.py:345-346
        losses = self.loss(out, gt)
        if not self._reducer.attached():
            self._reducer.attach()
            self._reducer.flat().zero_()
        # gradient-accumulation fusion for the [N, C] latent tables: their scatter-add kernels write straight into the
        # flat buffer's views (fields._direct_grad), which the optimiser kernel cleared
        for p in self.params:
            p._spf_direct_grad = True
        # every p.grad is a view of the flat buffer: autograd accumulates into it in place.  It is all-zero here: the
        # optimiser kernel clears it in the same pass that consumes it (zero_grad, train.py:355).
        hook = self._early_n > 0 and not self._diag_skip_reduce and self._reducer.attached()
        if hook:
            fields.GRAD_READY_HOOKS["color_latent"] = self._early_reduce
        side = self.wgrad_side and self._reducer.attached() and self.model.precision == "bf16"
        fields.wgrad_side_arm(side)
        try:
            losses["loss"].backward()
        finally:
            fields.GRAD_READY_HOOKS["color_latent"] = None
            fields.wgrad_side_arm(False)
        fields.wgrad_side_join()          # the colour field's weight gradients (side stream) are in the flat buffer
        self._allreduce_grads()           # N > 1: one NCCL all-reduce (sum) of the flat buffer, no packing
        # clip_grad_norm_(1.0) (train.py:360-361), the NaN / Inf guard (train.py:548-564: a non-finite global norm skips
        # the whole update, exactly like the reference's dropped gradients), Adam and zero_grad: two kernels, no host sync.
        self.opt.step(grad_scale=1.0 / self.world_size, zero_grad=True)
        return losses
