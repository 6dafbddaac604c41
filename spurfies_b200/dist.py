"""Ray-sharded data parallelism around the hot path (new functionality: the reference is single-GPU).

One process per GPU.  The step's rays are the independent units: rank r takes a contiguous slice of the ray batch;
the neural points, voxel grid, latent tables, MLPs and beta are replicated.  The ONLY exchange per training step is
one all-reduce (NCCL over NVLink on the GPU box; gloo in the CPU tests) of a single flat fp32 buffer holding the
gradients of every trainable tensor.  Gradient clipping (train.py:360-361) needs the global norm, so it runs after
the all-reduce; every rank then takes the identical Adam step.  Eval / meshing shard rays / grid slabs with no
collective at all (`shard_range` only).

Nothing here touches the kernels: it is host plumbing, and works on CPU tensors so that world_size-2 gloo tests can
cover it without a GPU.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n units for `rank`; the first n % world ranks get one extra unit."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(int(n), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(batch: dict, rank: int, world: int, ray_keys=("uv",), dim: int = 1) -> dict:
    """Slice the per-ray entries of a model input dict (uv [1,R,2]) for this rank; everything else is shared."""
    out = dict(batch)
    for k in ray_keys:
        lo, hi = shard_range(batch[k].shape[dim], rank, world)
        out[k] = batch[k].narrow(dim, lo, hi - lo)
    return out


class FlatGradReducer:
    """Flatten -> one all-reduce(sum) -> scale by 1/world -> write back.  The flat buffer is persistent (its address
    is stable, so the reduction can be captured in a CUDA graph together with the rest of the step)."""

    def __init__(self, params: Iterable[torch.Tensor], world_size: int, group: Optional[dist.ProcessGroup] = None):
        self.params: List[torch.Tensor] = list(params)
        self.world_size = int(world_size)
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self._flat: Optional[torch.Tensor] = None

    @property
    def bytes_per_step(self) -> int:
        return 4 * self.numel

    def flat(self) -> torch.Tensor:
        if self._flat is None:
            p0 = self.params[0]
            self._flat = torch.zeros(self.numel, dtype=torch.float32, device=p0.device)
        return self._flat

    def attach(self) -> torch.Tensor:
        """Make every `p.grad` a view into the flat buffer.  Autograd then accumulates straight into it (in place), so
        the all-reduce, the global-norm clip and the NaN guard are single passes over one tensor with no packing.
        The owner must clear gradients with `zero()` (not `zero_grad(set_to_none=True)`, which would drop the views)."""
        flat = self.flat()
        off = 0
        for p in self.params:
            k = p.numel()
            p.grad = flat[off:off + k].view_as(p)
            off += k
        return flat

    def attached(self) -> bool:
        if self._flat is None:
            return False
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self._flat.data_ptr() + 4 * off or not p.grad.is_contiguous():
                return False
            off += p.numel()
        return True

    def zero(self) -> None:
        self.flat().zero_()

    def reduce(self) -> None:
        """Average `p.grad` over the ranks in place (a missing grad counts as zero).  No-op for world_size 1."""
        if self.world_size <= 1:
            return
        flat = self.flat()
        if self.attached():
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.mul_(1.0 / self.world_size)
            return
        off = 0
        for p in self.params:
            k = p.numel()
            if p.grad is None:
                flat[off:off + k].zero_()
            else:
                flat[off:off + k].copy_(p.grad.reshape(-1))
            off += k
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.mul_(1.0 / self.world_size)
        off = 0
        for p in self.params:
            k = p.numel()
            g = flat[off:off + k].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += k


def max_over_ranks(values: List[float], device) -> List[float]:
    """Timing reduction used by bench.py: every multi-GPU number is the MAX over ranks."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]
