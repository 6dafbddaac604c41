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


def global_count_scales(local_counts: torch.Tensor, world: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Loss terms that are MEANS over a data-dependent count (valid samples, hit rays, crossing rays) must be normalised
    by the GLOBAL count for a ray-sharded step to equal the one-big-batch step (sum of numerators / global
    counts; reference means: pointneus_disent.py:765-780 ``F.l1_loss(..., reduction="mean")`` over all hit rays,
    loss.py:34-40 over all valid samples, feat_utils.py:446-451 over all crossing rays).  Given this rank's counts
    [k], returns the factors  local * world / global  that turn each rank-local mean into this rank's share of the global
    mean AFTER the gradient average over ranks (sum_r share_r / world == global mean).  One tiny all-reduce, no host
    sync, CUDA-graph capturable.  A term nobody contributes to (global count 0) keeps factor 1."""
    c = local_counts.detach().to(torch.float32)
    if world <= 1:
        return torch.ones_like(c)
    tot = c.clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    return torch.where(tot > 0, c * float(world) / tot.clamp(min=1.0), torch.ones_like(c))


def flat_offsets(params: Iterable[torch.Tensor], align: int = 1) -> Tuple[List[int], int]:
    """Element offsets of each tensor inside a flat buffer whose segments start on multiples of `align` elements
    (align = 4: 16-byte boundaries for the fused optimiser's 128-bit accesses) and the padded total."""
    offs, off = [], 0
    for p in params:
        offs.append(off)
        off += (p.numel() + align - 1) // align * align
    return offs, off


class FlatGradReducer:
    """Flatten -> one all-reduce(sum) -> scale by 1/world -> write back.  The flat buffer is persistent (its address
    is stable, so the reduction can be captured in a CUDA graph together with the rest of the step)."""

    def __init__(self, params: Iterable[torch.Tensor], world_size: int, group: Optional[dist.ProcessGroup] = None,
                 align: int = 1, bf16_prefix: int = 0):
        """bf16_prefix (opt-in, default off): the first `bf16_prefix` elements of the flat buffer -- the [N, C] latent
        tables, which are the first parameters and 96 % of the bytes -- travel as bf16 (rounded once before the sum,
        the sum itself rounded per hop by the collective); the rest (MLP weights, beta) stays fp32.  Meant for the
        bf16 precision mode, whose gradients already carry bf16-level noise; the fp32 mode keeps the exact reduction."""
        self.params: List[torch.Tensor] = list(params)
        self.world_size = int(world_size)
        self.group = group
        self.offsets, self.numel = flat_offsets(self.params, align)   # padding (if any) stays zero
        self._flat: Optional[torch.Tensor] = None
        self.bf16_prefix = max(0, min(int(bf16_prefix), self.numel))
        self._half: Optional[torch.Tensor] = None

    @property
    def bytes_per_step(self) -> int:
        return 4 * self.numel - 2 * self.bf16_prefix

    def _all_reduce_flat(self, flat: torch.Tensor) -> None:
        n = self.bf16_prefix
        if n == 0:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            return
        if self._half is None or self._half.device != flat.device:
            self._half = torch.empty(n, dtype=torch.bfloat16, device=flat.device)   # persistent: graph-capturable
        self._half.copy_(flat[:n])
        dist.all_reduce(self._half, op=dist.ReduceOp.SUM, group=self.group)
        if n < self.numel:
            dist.all_reduce(flat[n:], op=dist.ReduceOp.SUM, group=self.group)
        flat[:n].copy_(self._half)

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
        for p, off in zip(self.params, self.offsets):
            p.grad = flat[off:off + p.numel()].view_as(p)
        return flat

    def attached(self) -> bool:
        if self._flat is None:
            return False
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self._flat.data_ptr() + 4 * off or not p.grad.is_contiguous():
                return False
        return True

    def zero(self) -> None:
        self.flat().zero_()

    def reduce(self, average: bool = True) -> None:
        """Sum `p.grad` over the ranks in place and (average=True) divide by the world size; a missing grad counts as
        zero.  average=False leaves the sum: the fused optimiser folds 1/world into its clip coefficient.
        No-op for world_size 1."""
        if self.world_size <= 1:
            return
        flat = self.flat()
        if self.attached():
            self._all_reduce_flat(flat)
            if average:
                flat.mul_(1.0 / self.world_size)
            return
        for p, off in zip(self.params, self.offsets):
            k = p.numel()
            if p.grad is None:
                flat[off:off + k].zero_()
            else:
                flat[off:off + k].copy_(p.grad.reshape(-1))
        self._all_reduce_flat(flat)
        if average:
            flat.mul_(1.0 / self.world_size)
        for p, off in zip(self.params, self.offsets):
            k = p.numel()
            g = flat[off:off + k].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)


def max_over_ranks(values: List[float], device) -> List[float]:
    """Timing reduction used by bench.py: every multi-GPU number is the MAX over ranks."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]
