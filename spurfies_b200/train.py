"""One optimisation step of the reference trainer around the hot path (spurfies/train.py:330-397): forward,
VolSDFLoss, backward, clip_grad_norm_(1.0), NaN guard, Adam (lr 5e-4 on every trainable tensor, train.py:168-189).

Data parallel (new functionality): one process per GPU, the step's rays are sharded across
ranks, the neural points / grid / latents / MLPs are replicated, and ONE NCCL all-reduce per step carries the
flat fp32 gradient buffer (latents N x 96 + colour MLP + radiance head + beta).  Clipping uses the global norm,
so it runs after the all-reduce; every rank then takes the identical Adam step.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from .dist import FlatGradReducer
from .model import PointVolSDF, VolSDFLoss


class TrainStep:
    def __init__(self, model: PointVolSDF, lr: float = 5.0e-4, grad_clip: float = 1.0, loss: Optional[VolSDFLoss] = None,
                 world_size: int = 1):
        self.model = model
        self.loss = loss or VolSDFLoss()
        for prm in list(model.F_geometry.parameters()) + list(model.T.parameters()):
            prm.requires_grad_(False)  # the local-prior SDF field is frozen (train.py:151-154)
        self.params = [p for p in model.parameters() if p.requires_grad]
        cuda = self.params[0].is_cuda
        self.opt = torch.optim.Adam(self.params, lr=lr, fused=cuda, capturable=cuda)
        self.grad_clip = grad_clip
        self.world_size = world_size
        self._reducer = FlatGradReducer(self.params, world_size)
        self._graph = None
        self._static = None
        self.graph_error = None

    def _allreduce_grads(self):
        self._reducer.reduce()

    # ------------------------------------------------------------------ CUDA-graph replay of the whole step
    def capture(self, batch, gt, rng) -> bool:
        """Capture forward + loss + backward + (all-reduce) + clip + Adam into one CUDA graph.  Possible because the
        bf16 step has no host synchronisation (slot counts stay on the device) and every big buffer is persistent.
        Inputs are copied into static tensors before each replay.  Returns False (and stays eager) if capture fails."""
        if self.model.precision != "bf16":
            self.graph_error = "exact (fp32) mode sizes its library wgrad GEMMs on the host"
            return False
        try:
            import gc
            self.model._last = None
            gc.collect()  # drop every reference to earlier eager steps' autograd graphs (see model.forward)
            self._static = ({k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()},
                            {k: v.clone() for k, v in gt.items()}, {k: v.clone() for k, v in rng.items()})
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._eager(*self._static)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._static_out = self._eager(*self._static)
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
        sb, sg, sr = self._static
        for k, v in batch.items():
            if torch.is_tensor(v):
                sb[k].copy_(v, non_blocking=True)
        for k, v in gt.items():
            sg[k].copy_(v, non_blocking=True)
        for k, v in rng.items():
            sr[k].copy_(v, non_blocking=True)
        self._graph.replay()
        return self._static_out

    def __call__(self, batch, gt, rng=None) -> Dict[str, torch.Tensor]:
        if self._graph is not None:
            return self.replay(batch, gt, rng)
        return self._eager(batch, gt, rng)

    def _eager(self, batch: Dict[str, torch.Tensor], gt: Dict[str, torch.Tensor], rng=None) -> Dict[str, torch.Tensor]:
        self.model.train()
        out = self.model(batch, fast=1, rng=rng, dense_outputs=True)  # fast=1: train.py:345-346
        losses = self.loss(out, gt)
        if not self._reducer.attached():
            self._reducer.attach()
        flat = self._reducer.flat()
        flat.zero_()                      # every p.grad is a view of `flat`: autograd accumulates into it in place
        losses["loss"].backward()
        self._allreduce_grads()           # N > 1: one NCCL all-reduce of `flat`, no packing
        # clip_grad_norm_(1.0) (train.py:360-361) and the NaN / Inf guard (train.py:548-564) share ONE reduction: the
        # global norm is finite iff every gradient entry is.  When it is not, the gradients are zeroed (the reference
        # drops them and skips the update; here Adam still decays its moments on such a step).  No host sync.
        total = torch.linalg.vector_norm(flat, 2.0)
        coef = torch.clamp(self.grad_clip / (total + 1e-6), max=1.0) if self.grad_clip > 0 else torch.ones_like(total)
        coef = torch.where(torch.isfinite(total), coef, torch.zeros_like(coef))
        flat.mul_(coef)
        torch.nan_to_num_(flat, nan=0.0, posinf=0.0, neginf=0.0)
        self.opt.step()
        return losses
