"""Host-side wrappers (torch.autograd.Function) around the field / compositing kernels of libspurfies_b200.so.

PyTorch is used for device memory, streams and autograd bookkeeping only; every tensor op on the hot path is a
hand-written kernel reached through the C ABI (include/spurfies_b200.h).  There is no library GEMM: the weight-gradient
products dW = dZ^T @ A run in the split-K tcgen05 kernel (spf_wgrad_tc_multi) in the tensor-core mode and in the split-K
fp32 FFMA kernel (spf_wgrad_f32) in the exact mode.

Layout: a *slot* is one query position (ray sample or point).  ``pidx`` [n, K] holds its neighbours sorted by
(d^2, id), -1 padded.  Valid slots (>= 1 neighbour) are compacted on the device into ``list`` / ``count``; per-slot
outputs (sdf, grad, hbar, rgb) are indexed by slot, per-pair saved tensors by compact row ``v*K + k``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import (ColorWeightsF32, ColorWeightsTC, GeoWeightsF32, GeoWeightsTC, HeadWeightsF32, HeadWeightsTC, call, ptr,
                   stream)
from .packing import image_bytes, pack_flush, pack_sw128, pack_sw128_dev

K_NEIGH = 8
ROW_PAD = 128  # saved per-pair tensors are written in whole tiles


class Arena:
    """Persistent device buffers for the big per-step intermediates (saved activations, Jacobian rows, dZ rows).
    They are sized for the worst case (every slot valid) and would otherwise be cudaMalloc'ed and freed every step
    (several GB): the caching allocator then dominates the step.  Everything is stream-ordered, so reusing the same
    storage step after step is safe -- as long as ONE forward per (model, query tag) is in flight between a forward and
    its backward.  That is enforced, not assumed: buffer names carry a per-model prefix (two models never share
    storage), every SlotSet claims a generation of its tag, and the backward passes call ``SlotSet.check_live()``, which
    raises if a later forward of the same model re-issued the buffers holding their saved activations."""
    _bufs = {}
    _gen = {}

    @classmethod
    def get(cls, name: str, shape, dtype, device) -> torch.Tensor:
        key = (name, str(device))
        numel = 1
        for d in shape:
            numel *= int(d)
        nbytes = numel * torch.empty(0, dtype=dtype).element_size()
        buf = cls._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
            cls._bufs[key] = buf
        return buf[:nbytes].view(dtype).view(*shape)

    @classmethod
    def claim(cls, tag: str, device) -> int:
        key = (tag, str(device))
        cls._gen[key] = cls._gen.get(key, 0) + 1
        return cls._gen[key]

    @classmethod
    def current(cls, tag: str, device) -> int:
        return cls._gen.get((tag, str(device)), 0)

    @classmethod
    def clear(cls):
        cls._bufs.clear()


class SlotSet:
    """Compacted list of the valid slots of one query (utils.py:90-113 glue, without the host sync)."""

    def __init__(self, pidx: torch.Tensor, tag: str = "q", owner: str = ""):
        assert pidx.dtype == torch.int32 and pidx.is_contiguous()
        self.owner = owner       # per-model prefix of every arena buffer this query uses
        self.tag = owner + tag   # names this query's arena buffers ("fine", "points", "pseudo", ...)
        self.gen = Arena.claim(self.tag, pidx.device)
        self.K = pidx.shape[-1]
        self.pidx = pidx.view(-1, self.K)
        self.n = self.pidx.shape[0]
        dev = pidx.device
        self.list = torch.empty(max(self.n, 1), dtype=torch.int32, device=dev)
        self.count = torch.empty(1, dtype=torch.int32, device=dev)   # always written by spf_compact_valid
        ws = Arena.get("compact_ws", (_lib.lib.spf_compact_workspace_bytes(self.n),), torch.uint8, dev)
        call("spf_compact_valid", ptr(self.pidx), self.n, self.K, ptr(self.list), ptr(self.count), ptr(ws), ws.numel(),
             stream())
        self._V: Optional[int] = None

    @property
    def V(self) -> int:
        """Number of valid slots ON THE HOST (one sync).  Only the reference's ragged return contracts and the exact
        (fp32) mode's library wgrad GEMMs need it; the bf16 training step never does (CUDA-graph capturable)."""
        if self._V is None:
            self._V = int(self.count.item())
        return self._V

    def valid_mask(self) -> torch.Tensor:
        return self.pidx[:, 0] >= 0

    def check_live(self) -> None:
        """Called by every backward that reads activations saved in the arena under this query's tag."""
        cur = Arena.current(self.tag, self.pidx.device)
        if cur != self.gen:
            raise RuntimeError(
                f"spurfies_b200: the saved activations of query '{self.tag}' (generation {self.gen}) were overwritten by a "
                f"later forward of the same model (generation {cur}) before this backward ran.  Run backward() before the "
                "next forward / render / pseudo_sdf call of this model, or use a second model instance.")

    def rows_alloc(self, per_slot: int) -> int:
        return self.n * per_slot + ROW_PAD


# ------------------------------------------------------------------------------------------------ weight packs
class GeoPack:
    """Frozen geometry MLP (F_geometry + T, pointneus_disent.py:86-98) packed for the kernels; rebuilt only when
    a parameter's version counter changes."""

    def __init__(self):
        self._key = None

    def get(self, F_geometry, T):
        lin = [m for m in F_geometry if isinstance(m, torch.nn.Linear)]
        prm = [p for m in lin for p in (m.weight, m.bias)] + [T[0].weight, T[0].bias]
        key = tuple((p.data_ptr(), p._version) for p in prm)
        if key != self._key:
            with torch.no_grad():
                W = [m.weight.detach().float().contiguous() for m in lin]
                b = [m.bias.detach().float().contiguous() for m in lin]
                tw, tb = T[0].weight.detach().double(), T[0].bias.detach().double()
                # fold F_geometry.8 and T (both linear, no activation in between)
                v5 = (tw @ W[4].double()).reshape(-1).float().contiguous()
                c5 = float((tw @ b[4].double() + tb).reshape(()))
                self.W, self.b = W, b
                self.Wt = [w.t().contiguous() for w in W[:4]]
                self.v5, self.c5 = v5, c5
                s = GeoWeightsF32()
                s.w1t, s.b1 = self.Wt[0].data_ptr(), b[0].data_ptr()
                s.w2t, s.b2 = self.Wt[1].data_ptr(), b[1].data_ptr()
                s.w3t, s.b3 = self.Wt[2].data_ptr(), b[2].data_ptr()
                s.w4t, s.b4 = self.Wt[3].data_ptr(), b[3].data_ptr()
                s.v5, s.c5 = v5.data_ptr(), c5
                s.w1, s.w2, s.w3, s.w4 = (W[i].data_ptr() for i in range(4))
                self.f32 = s
                # tensor-core images (mlp_tc2.cu): x - p enters twice (fp16 hi + lo) against the same weights
                W1ext = torch.cat([W[0][:, :32], W[0][:, 32:35], W[0][:, 32:35]], dim=1)
                f16 = torch.float16   # fp16 operands for the forward layers and the d sdf / d input chain (csrc/umma.cuh)
                self.tc_imgs = [pack_sw128(W1ext, dtype=f16), pack_sw128(W[1], dtype=f16), pack_sw128(W[2], dtype=f16),
                                pack_sw128(W[3], dtype=f16),
                                pack_sw128(W[3].t(), dtype=f16), pack_sw128(W[2].t(), dtype=f16),
                                pack_sw128(W[1].t(), dtype=f16), pack_sw128(W[0].t(), n_pad=48, dtype=f16)]
                t = GeoWeightsTC()
                (t.w1p, t.w2p, t.w3p, t.w4p, t.w4tp, t.w3tp, t.w2tp, t.w1tp) = (i.data_ptr() for i in self.tc_imgs)
                t.b1, t.b2, t.b3, t.b4 = (b[i].data_ptr() for i in range(4))
                t.v5, t.c5 = v5.data_ptr(), c5
                self.tc = t
            self._key = key
        return self


PRECISION = {"mode": "fp32"}  # "fp32": exact SIMT kernels (1e-4); "bf16": tcgen05 tensor-core kernels (2e-2)


def set_precision(mode: str):
    assert mode in ("fp32", "bf16")
    PRECISION["mode"] = mode


def geo_sdf_raw(pack: GeoPack, slots: SlotSet, x, pts, feat_g, rbf, want_grad, want_jw, fill=1000.0):
    n = slots.n
    dev = x.device
    sdf = torch.full((n,), fill, dtype=torch.float32, device=dev)
    grad = torch.zeros(n, 3, dtype=torch.float32, device=dev) if want_grad else None
    jw = Arena.get(slots.tag + ".jw", (slots.rows_alloc(slots.K), 32), torch.float32, dev) if want_jw else None
    if PRECISION["mode"] == "bf16":
        call("spf_sdf_fwd_tc", C.byref(pack.tc), ptr(slots.list), ptr(slots.count), n, ptr(x), ptr(slots.pidx), slots.K,
             ptr(pts), ptr(feat_g), float(rbf), ptr(sdf), ptr(grad), ptr(jw), stream())
    else:
        call("spf_sdf_fwd_f32", C.byref(pack.f32), ptr(slots.list), ptr(slots.count), n, ptr(x), ptr(slots.pidx), slots.K,
             ptr(pts), ptr(feat_g), float(rbf), ptr(sdf), ptr(grad), ptr(jw), stream())
    return sdf, grad, jw


def _direct_grad(param) -> Optional[torch.Tensor]:
    """Gradient-accumulation fusion (opt-in, set by TrainStep on the latent tables once their .grad is a view of the flat,
    optimiser-cleared gradient buffer): the scatter-add kernels then accumulate straight into ``param.grad`` and the
    Function returns None for that input -- no zero-filled temporary, no AccumulateGrad pass over the [N, C] table."""
    if getattr(param, "_spf_direct_grad", False) and param.grad is not None and param.grad.is_contiguous():
        return param.grad
    return None


class GeoSDF(torch.autograd.Function):
    """sdf[slot] = sum_k w_k T(F_geometry([g_k | x - p_k])) / sum_k w_k and d sdf / d x  (pointneus_disent.py:241-247,
    300-323).  Gradients: latents (scatter-add of the saved Jacobian rows) and, when x requires grad (pseudo-point
    path, pointneus_disent.py:771-780), x."""

    @staticmethod
    def forward(ctx, feat_g, x, slots: SlotSet, pack: GeoPack, pts, rbf, want_grad):
        feat_needs, x_needs = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        x = x.detach().contiguous()
        sdf, grad, jw = geo_sdf_raw(pack, slots, x, pts, feat_g.detach(), rbf, want_grad or x_needs, feat_needs)
        ctx.slots, ctx.jw, ctx.grad = slots, jw, grad
        ctx.feat_shape = feat_g.shape
        ctx.direct = _direct_grad(feat_g)
        ctx.x_needs = x_needs
        if grad is None:
            grad = torch.zeros(0, 3, device=x.device)
        ctx.mark_non_differentiable(grad)
        return sdf, grad

    @staticmethod
    def backward(ctx, d_sdf, _):
        slots = ctx.slots
        if ctx.jw is not None:
            slots.check_live()
        gfeat = None
        d_sdf = d_sdf.contiguous()
        if ctx.jw is not None:
            target = ctx.direct
            if target is None:
                target = gfeat = torch.zeros(ctx.feat_shape, dtype=torch.float32, device=d_sdf.device)
            call("spf_sdf_bwd", ptr(slots.list), ptr(slots.count), slots.n, ptr(slots.pidx), slots.K, ptr(ctx.jw),
                 ptr(d_sdf), ptr(target), stream())
        dx = None
        if ctx.x_needs:
            dx = torch.where(slots.valid_mask()[:, None], d_sdf[:, None] * ctx.grad, torch.zeros_like(ctx.grad))
        return gfeat, dx, None, None, None, None, None


def _color_struct(W, b):
    s = ColorWeightsF32()
    Wt = [w.t().contiguous() for w in W]
    s.w1t, s.b1, s.w2t, s.b2, s.w3t, s.b3 = (Wt[0].data_ptr(), b[0].data_ptr(), Wt[1].data_ptr(), b[1].data_ptr(),
                                             Wt[2].data_ptr(), b[2].data_ptr())
    s.w1, s.w2, s.w3 = W[0].data_ptr(), W[1].data_ptr(), W[2].data_ptr()
    return s, Wt


def _img(name, dev, nbytes):
    return Arena.get("img." + name, (nbytes,), torch.uint8, dev)


def _color_struct_tc(W, b, owner=""):
    """bf16 weight images for k_color_fwd_tc / k_color_bwd_tc (input columns permuted to [c (64) | PE6 (39)]),
    packed on the device by one spf_pack_sw128_batch launch into persistent (per-model) buffers."""
    dev = W[0].device
    jobs = []
    w1p = _img(owner + "c.w1p", dev, 65536)
    pack_sw128_dev(w1p, W[0], 256, 64, col_off=39, batch=jobs, f16=True)                       # k-block 0: latent columns
    pack_sw128_dev(w1p, W[0], 256, 39, col_off=0, row_off_bytes=32768, batch=jobs, f16=True)   # k-block 1: PE6 columns
    imgs = [w1p]
    for nm, w, tr in (("c.w2p", W[1], False), ("c.w3p", W[2], False), ("c.w3tp", W[2], True), ("c.w2tp", W[1], True)):
        im = _img(owner + nm, dev, 131072)
        pack_sw128_dev(im, w, 256, 256, transpose=tr, batch=jobs, f16=True)   # every operand of the tensor-core mode is fp16
        imgs.append(im)
    w1ftp = _img(owner + "c.w1ftp", dev, 32768)
    pack_sw128_dev(w1ftp, W[0], 64, 256, transpose=True, col_off=39, batch=jobs, f16=True)     # (W1[:, 39:103])^T : [64][256]
    imgs.append(w1ftp)
    pack_flush(jobs)   # one launch
    s = ColorWeightsTC()
    s.w1p, s.w2p, s.w3p, s.w3tp, s.w2tp, s.w1ftp = (i.data_ptr() for i in imgs)
    s.b1, s.b2, s.b3 = (v.data_ptr() for v in b)
    return s, imgs


class _ZeroPool:
    """One zero-filled fp32 buffer per backward call, handed out in 16-byte aligned pieces (the split-K weight-gradient
    kernel accumulates with atomics, so its outputs start at zero: one fill instead of one per tensor)."""

    def __init__(self, numel: int, device):
        self.buf = torch.zeros(numel, dtype=torch.float32, device=device)
        self.off = 0

    def take(self, *shape) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        assert self.off + n <= self.buf.numel(), "zero pool exhausted"
        out = self.buf[self.off:self.off + n].view(*shape)
        self.off += (n + 3) // 4 * 4
        return out


_GS_SCRATCH = {}


def grad_scale(upstream: torch.Tensor, target: float = 16.0) -> torch.Tensor:
    """Device tensor [S, 1/S] for the tensor-core mode's fp16 gradient chain: S = the power of two that brings the largest
    upstream gradient entry just below `target` (2^12 of fp16 headroom above for growth through the layers, 2^28 of range
    below), computed on the device by one launch (spf_grad_scale: no host sync, CUDA-graph capturable).  Powers of two
    make scaling / unscaling exact.  (The kernel's two scratch words are per device: calls must be stream-ordered with
    respect to each other, as the step's two calls are.)"""
    x = upstream.detach()
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.float().contiguous()
    dev = x.device
    scratch = _GS_SCRATCH.get(dev)
    if scratch is None:
        scratch = _GS_SCRATCH[dev] = torch.zeros(2, dtype=torch.int32, device=dev)   # kept zero by the kernel
    out = torch.empty(2, dtype=torch.float32, device=dev)
    if x.numel() == 0:
        return out.fill_(1.0)
    call("spf_grad_scale", ptr(x), x.numel(), float(target), ptr(out), ptr(scratch), stream())
    return out


def _wgrad_multi(jobs, slots, rows_per_unit, pool: _ZeroPool, targets=None, gscale=None):
    """[(dz, act, lda, N, want_db[, fmt])] -> [(dW [256,N], db [256] or None)]: all products in ONE spf_wgrad_tc_multi launch
    (every operand in the tile layout, lda a multiple of 64); no host sync.  fmt (default 0 = both fp16): bit 0 = dz is
    bf16, bit 1 = act is bf16.  `gscale` = the step's [S, 1/S] (grad_scale): the gradient tiles carry S, the outputs do not.  ``targets[i] = (dW_grad, db_grad)`` (either may be
    None) makes job i accumulate straight into those gradient buffers (see _direct_grad); None is then returned for them."""
    arr = (_lib.WgradJob * len(jobs))()
    out = []
    for i, job in enumerate(jobs):
        dz, act, lda, N, want_db = job[:5]
        arr[i].fmt = job[5] if len(job) > 5 else 0
        tW, tb = targets[i] if targets is not None else (None, None)
        dW = tW if tW is not None else pool.take(256, N)
        db = (tb if tb is not None else pool.take(256)) if want_db else None
        arr[i].dz, arr[i].act, arr[i].dW = dz.data_ptr(), act.data_ptr(), dW.data_ptr()
        arr[i].db = db.data_ptr() if db is not None else None
        arr[i].lda, arr[i].N = int(lda), int(N)
        out.append((None if tW is not None else dW, None if (tb is not None or db is None) else db))
    call("spf_wgrad_tc_multi", C.cast(arr, C.c_void_p), len(jobs), ptr(slots.count), int(rows_per_unit), slots.n,
         ptr(gscale), stream())
    return out


# Side stream for the colour field's weight-gradient launch (armed by train.TrainStep for the duration of one backward,
# joined by it before the gradient exchange / optimiser).
WGRAD_SIDE = {"armed": False, "stream": None, "pending": False, "keep": []}


def wgrad_side_arm(on: bool) -> None:
    if on and WGRAD_SIDE["stream"] is None:
        WGRAD_SIDE["stream"] = torch.cuda.Stream()
    WGRAD_SIDE["armed"] = bool(on)


def wgrad_side_join() -> None:
    """The current stream waits for the side-stream weight-gradient launch of this backward (no host sync)."""
    if WGRAD_SIDE["pending"]:
        torch.cuda.current_stream().wait_stream(WGRAD_SIDE["stream"])
        WGRAD_SIDE["pending"] = False
    # only now may the caching allocator hand the launch's main-stream inputs to somebody else (see ColorField.backward)
    WGRAD_SIDE["keep"].clear()


# callbacks fired inside the backward as soon as a parameter's gradient is final (set by train.TrainStep for the
# duration of one backward; None = nobody listens)
GRAD_READY_HOOKS = {"color_latent": None}


class ColorField(torch.autograd.Function):
    """hbar[slot] = sum_k w_k/norm * h3_k with h3 = first three layers of F_color on [PE6(x-p_k) | c_k]
    (pointneus_disent.py:325-336; F_color.6 is applied per sample in RadianceHead)."""

    @staticmethod
    def forward(ctx, feat_c, W1, b1, W2, b2, W3, b3, x, slots: SlotSet, pts, rbf):
        dev = x.device
        tcm = PRECISION["mode"] == "bf16"
        W = [w.detach().float().contiguous() for w in (W1, W2, W3)]
        b = [v.detach().float().contiguous() for v in (b1, b2, b3)]
        s, keep = _color_struct_tc(W, b, slots.owner) if tcm else _color_struct(W, b)
        n, K = slots.n, slots.K
        tg = slots.tag
        hbar = Arena.get(tg + ".hbar", (n, 256), torch.float32, dev)  # read back only at valid slots
        need = any(ctx.needs_input_grad[:7])
        rows = slots.rows_alloc(K)
        in0 = h1 = h2 = m3 = wn = None
        if need:
            adt = torch.float16 if tcm else torch.float32
            in0 = Arena.get(tg + ".in0", (rows, 128 if tcm else 104), adt, dev)  # tc: tile layout, 2 k-blocks
            h1 = Arena.get(tg + ".h1", (rows, 256), adt, dev)
            h2 = Arena.get(tg + ".h2", (rows, 256), adt, dev)
            m3 = Arena.get(tg + ".m3", (rows, 24), torch.int32, dev)  # LeakyReLU sign words of z1..z3 (8 per layer)
            wn = Arena.get(tg + ".wn", (rows,), torch.float32, dev)
        if tcm:
            # the kernel also leaves a bf16 copy of hbar indexed by COMPACT slot in the tile layout: the radiance head
            # bulk-copies it as its A operand (and its F_color.6 weight gradient reads it) instead of gathering fp32 rows
            hb = Arena.get(tg + ".hhb", (slots.rows_alloc(1), 256), torch.float16, dev)
            call("spf_color_fwd_tc", C.byref(s), ptr(slots.list), ptr(slots.count), n, ptr(x.contiguous()), ptr(slots.pidx),
                 K, ptr(pts), ptr(feat_c.detach()), float(rbf), ptr(hbar), ptr(in0), ptr(h1), ptr(h2), ptr(m3), ptr(wn),
                 ptr(hb), stream())
            slots.hb_compact = (hb, hbar.data_ptr(), hbar._version)
        else:
            call("spf_color_fwd_f32", C.byref(s), ptr(slots.list), ptr(slots.count), n, ptr(x.contiguous()),
                 ptr(slots.pidx), K, ptr(pts), ptr(feat_c.detach()), float(rbf), ptr(hbar), ptr(in0), ptr(h1), ptr(h2),
                 ptr(m3), ptr(wn), stream())
        ctx.slots, ctx.saved_t, ctx.tcm = slots, (s, keep, W, b, in0, h1, h2, m3, wn), tcm
        ctx.feat_shape = feat_c.shape
        ctx.direct = _direct_grad(feat_c)
        ctx.direct_w = [_direct_grad(p) for p in (W1, b1, W2, b2, W3, b3)] if tcm else [None] * 6
        return hbar

    @staticmethod
    def backward(ctx, d_hbar):
        slots, tcm = ctx.slots, ctx.tcm
        slots.check_live()
        s, keep, W, b, in0, h1, h2, m3, wn = ctx.saved_t
        dev = d_hbar.device
        rows = slots.rows_alloc(slots.K)
        adt = torch.float16 if tcm else torch.float32
        tg = slots.tag
        dz1 = Arena.get(tg + ".dz1", (rows, 256), adt, dev)
        dz2 = Arena.get(tg + ".dz2", (rows, 256), adt, dev)
        dz3 = Arena.get(tg + ".dz3", (rows, 256), adt, dev)
        gfeat = ctx.direct if ctx.direct is not None else torch.zeros(ctx.feat_shape, dtype=torch.float32, device=dev)
        if tcm:
            # the radiance head may have left its gradient as bf16 by compact sample row in the tile layout (see
            # RadianceHead.backward); `d_hbar` is then only the placeholder autograd carried here
            cached = getattr(slots, "d_hb_compact", None)
            compact = cached is not None and cached[1] == d_hbar.data_ptr()
            # the gradient chain is fp16 scaled by the step's power of two S: the head's (its compact gradient already
            # carries it) or, for a stand-alone call, one derived from this upstream gradient
            gscale = cached[2] if compact else grad_scale(d_hbar)
            call("spf_color_bwd_tc", C.byref(s), ptr(slots.list), ptr(slots.count), slots.n, ptr(slots.pidx), slots.K,
                 None if compact else ptr(d_hbar.contiguous()), ptr(h1), ptr(h2), ptr(m3), ptr(wn), ptr(dz1), ptr(dz2),
                 ptr(dz3), ptr(gfeat), ptr(cached[0]) if compact else None, ptr(gscale), stream())
            slots.d_hb_compact = None
        else:
            call("spf_color_bwd_f32", C.byref(s), ptr(slots.list), ptr(slots.count), slots.n, ptr(slots.pidx), slots.K,
                 ptr(d_hbar.contiguous()), ptr(h1), ptr(h2), ptr(m3), ptr(wn), ptr(dz1), ptr(dz2), ptr(dz3), ptr(gfeat),
                 stream())
        if tcm:  # hand-written split-K tcgen05 wgrad, row count read on the device
            tw = ctx.direct_w   # W1's gradient needs its columns permuted back: through the pool; the rest may go direct

            def wgrad():
                pool = _ZeroPool(2 * 256 * 256 + 256 * 112 + 3 * 256, dev)
                (dW3, db3), (dW2, db2), (dW1p, db1) = _wgrad_multi(
                    [(dz3, h2, 256, 256, True), (dz2, h1, 256, 256, True), (dz1, in0, 128, 112, True)], slots, slots.K, pool,
                    targets=[(tw[4], tw[5]), (tw[2], tw[3]), (None, tw[1])], gscale=gscale)
                return torch.cat([dW1p[:, 64:103], dW1p[:, :64]], dim=1), db1, dW2, db2, dW3, db3

            side = WGRAD_SIDE["stream"] if (WGRAD_SIDE["armed"] and all(t is not None for t in tw)) else None
            if side is None:
                dW1, db1, dW2, db2, dW3, db3 = wgrad()
            else:
                # every gradient of this launch lands in the trainer's flat buffer, so nothing downstream of this node
                # reads it before the optimiser: run it on a side stream, under the geometry backward / regulariser /
                # pseudo-point backward that follow on this one (small latency-bound kernels that fit next to the
                # HBM-bound split-K kernel); train.TrainStep joins the stream before it touches the gradients
                # Its inputs were allocated on THIS stream: a block freed when this node returns (the [S, 1/S] pair, the
                # slot set) could be re-issued to a later main-stream kernel while the side launch still reads it, so
                # they stay referenced until the join.
                WGRAD_SIDE["keep"].append((gscale, slots, dz1, dz2, dz3, h1, h2, in0))
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    dW1, db1, dW2, db2, dW3, db3 = wgrad()
                    tw[0].add_(dW1)
                    dW1 = None
                WGRAD_SIDE["pending"] = True
        else:    # exact mode: fp32 FFMA split-K kernel over the compact pair rows (device-side row count, no library GEMM)
            dW3, db3 = _wgrad_f32(dz3, h2, 256, slots, slots.K)
            dW2, db2 = _wgrad_f32(dz2, h1, 256, slots, slots.K)
            dW1, db1 = _wgrad_f32(dz1, in0, 103, slots, slots.K)
        # The colour latents' gradient has been complete since the dgrad kernel (this is their only consumer): a
        # data-parallel trainer may start reducing it now, under the geometry backward and the regulariser that follow
        # (train.py).  Fired AFTER the weight-gradient launch: that kernel runs at 90 % of the HBM peak and NCCL's copy
        # kernels next to it cost more than they hide (8 GPUs: k_wgrad_multi 0.72 -> 0.96 ms, step 4.30 -> 4.57 ms).
        if ctx.direct is not None and GRAD_READY_HOOKS["color_latent"] is not None:
            GRAD_READY_HOOKS["color_latent"]()
        return (None if ctx.direct is not None else gfeat), dW1, db1, dW2, db2, dW3, db3, None, None, None, None


def _wgrad_f32(dz, act, N, slots, rows_per_unit, want_db=True, M=256, idx=None, idx_div=1):
    """dW [M,N] (+ db [M]) of the exact mode through spf_wgrad_f32: fp32 FFMA split-K, row count read on the device."""
    dev = dz.device
    dW = torch.zeros(M, N, dtype=torch.float32, device=dev)
    db = torch.zeros(M, dtype=torch.float32, device=dev) if want_db else None
    call("spf_wgrad_f32", ptr(dz), int(dz.stride(0)), int(M), ptr(act), int(act.stride(0)), int(N), ptr(idx), int(idx_div),
         ptr(slots.count), int(rows_per_unit), slots.n, ptr(dW), ptr(db), stream())
    return dW, db


def positional_encoding(x: torch.Tensor, multires: int) -> torch.Tensor:
    """embedder.py:10-36 (host-side copy, only used to assemble the wgrad operand of R.0)."""
    out = [x]
    for l in range(multires):
        out += [torch.sin(x * float(2 ** l)), torch.cos(x * float(2 ** l))]
    return torch.cat(out, -1)


class RadianceHead(torch.autograd.Function):
    """rgb[slot] = sigmoid(R([PE3(dir) | F_color.6(hbar)]))  (pointneus_disent.py:83, 100-107, 338-346)."""

    @staticmethod
    def forward(ctx, hbar, W4, b4, R1, rb1, R2, rb2, R3, rb3, dirs, slots: SlotSet, Smax):
        dev = hbar.device
        tcm = PRECISION["mode"] == "bf16"
        W = [w.detach().float().contiguous() for w in (W4, R1, R2, R3)]
        b = [v.detach().float().contiguous() for v in (b4, rb1, rb2, rb3)]
        n = slots.n
        tg = slots.tag
        rgb = torch.zeros(n, 3, dtype=torch.float32, device=dev)
        rows = slots.rows_alloc(1)
        need = any(ctx.needs_input_grad[:9])
        hbar_c = hbar.detach().contiguous()
        if tcm:
            imgs, jobs = [], []
            for nm, w, N_, K_, tr, npad, co in (("h.w4p", W[0], 256, 256, False, 256, 0), ("h.r1fp", W[1], 256, 256, False, 256, 21),
                                               ("h.r2p", W[2], 256, 256, False, 256, 0), ("h.r3p", W[3], 3, 256, False, 32, 0),
                                               ("h.r3tp", W[3], 256, 3, True, 256, 0), ("h.r2tp", W[2], 256, 256, True, 256, 0),
                                               ("h.r1ftp", W[1], 256, 256, True, 256, 21), ("h.w4tp", W[0], 256, 256, True, 256, 0)):
                im = _img(slots.owner + nm, dev, image_bytes(npad, K_))
                pack_sw128_dev(im, w, N_, K_, transpose=tr, n_pad=npad, col_off=co, batch=jobs, f16=True)
                imgs.append(im)
            pack_flush(jobs)
            s = HeadWeightsTC()
            s.w4p, s.r1fp, s.r2p, s.r3p, s.r3tp, s.r2tp, s.r1ftp, s.w4tp = (i.data_ptr() for i in imgs)
            s.b4, s.rb2, s.rb3 = b[0].data_ptr(), b[2].data_ptr(), b[3].data_ptr()
            # per-ray constant part of R.0: PE3(dir) columns + bias, kept in fp32
            zpe = Arena.get(tg + ".zpe", (dirs.shape[0], 256), torch.float32, dev)
            call("spf_head_zpe", ptr(dirs), ptr(W[1]), int(W[1].stride(0)), ptr(b[1]), int(dirs.shape[0]), ptr(zpe), stream())
            hb = f = a1 = a2 = pe = None
            # compact bf16 hbar left by ColorField.forward for exactly this tensor (same storage, not modified since)?
            cached = getattr(slots, "hb_compact", None)
            from_color = cached is not None and cached[1:] == (hbar.data_ptr(), hbar._version)
            if from_color:
                hb = cached[0]
            elif need:
                hb = Arena.get(tg + ".hhb", (rows, 256), torch.float16, dev)
            if need:
                f = Arena.get(tg + ".hf", (rows, 256), torch.float16, dev)
                a1 = Arena.get(tg + ".ha1", (rows, 256), torch.float16, dev)
                a2 = Arena.get(tg + ".ha2", (rows, 256), torch.float16, dev)
                pe = Arena.get(tg + ".hpe", (rows, 64), torch.float16, dev)   # tile layout, one k-block (32 columns used)
            call("spf_head_fwd_tc", C.byref(s), ptr(slots.list), ptr(slots.count), n, None if from_color else ptr(hbar_c),
                 ptr(zpe), ptr(dirs), int(Smax), ptr(rgb), ptr(hb), ptr(f), ptr(a1), ptr(a2), ptr(pe), stream())
            ctx.saved_t = (s, imgs, W, b, (hb, pe), rgb.detach(), f, a1, a2, dirs)
            ctx.from_color = from_color
            ctx.direct_w = [_direct_grad(p) for p in (W4, b4, R1, rb1, R2, rb2, R3, rb3)]
        else:
            Wt = [w.t().contiguous() for w in W]
            s = HeadWeightsF32()
            s.w4t, s.b4, s.r1t, s.rb1, s.r2t, s.rb2, s.r3t, s.rb3 = (Wt[0].data_ptr(), b[0].data_ptr(), Wt[1].data_ptr(),
                                                                      b[1].data_ptr(), Wt[2].data_ptr(), b[2].data_ptr(),
                                                                      Wt[3].data_ptr(), b[3].data_ptr())
            s.w4, s.r1, s.r2, s.r3 = (w.data_ptr() for w in W)
            f = a1 = a2 = None
            if need:
                f = Arena.get(tg + ".hf", (rows, 256), torch.float32, dev)
                a1 = Arena.get(tg + ".ha1", (rows, 256), torch.float32, dev)
                a2 = Arena.get(tg + ".ha2", (rows, 256), torch.float32, dev)
            call("spf_head_fwd_f32", C.byref(s), ptr(slots.list), ptr(slots.count), n, ptr(hbar_c), ptr(dirs), int(Smax),
                 ptr(rgb), ptr(f), ptr(a1), ptr(a2), stream())
            ctx.saved_t = (s, Wt, W, b, hbar_c, rgb.detach(), f, a1, a2, dirs)
        ctx.slots, ctx.Smax, ctx.tcm = slots, Smax, tcm
        return rgb

    @staticmethod
    def backward(ctx, d_rgb):
        slots, tcm = ctx.slots, ctx.tcm
        slots.check_live()
        s, keep, W, b, hb, rgb, f, a1, a2, dirs = ctx.saved_t
        dev = d_rgb.device
        n = slots.n
        tg = slots.tag
        rows = slots.rows_alloc(1)
        adt = torch.float16 if tcm else torch.float32
        d_hbar = Arena.get(tg + ".d_hbar", (n, 256), torch.float32, dev)  # consumed only at valid slots
        dzf = Arena.get(tg + ".hdzf", (rows, 256), adt, dev)
        dz1 = Arena.get(tg + ".hdz1", (rows, 256), adt, dev)
        dz2 = Arena.get(tg + ".hdz2", (rows, 256), adt, dev)
        if tcm:
            hb, pe = hb
            dz3 = Arena.get(tg + ".hdz3b", (rows, 64), torch.float16, dev)   # tile layout, one k-block (3 columns used)
            pool = _ZeroPool(3 * 256 * 256 + 256 * 32 + 256 * 16 + 3 * 256 + 4, dev)
            tw = ctx.direct_w   # R.0 (concatenated) and R.4 (transposed) go through the pool; the rest may go direct
            drb3 = tw[7] if tw[7] is not None else pool.take(3)
            # `slots.single_consumer` (set by PointVolSDF.forward): hbar feeds nothing but this head, so its gradient can
            # travel to ColorField.backward as bf16 by compact sample row in the tile layout (one coalesced bulk store per
            # tile here, half the bytes there); the fp32 `d_hbar` returned to autograd is then an unwritten placeholder
            d_hbc = None
            gscale = grad_scale(d_rgb)   # [S, 1/S]: the fp16 gradient chain of this step (csrc/mlp_tc2.cu: mask_pack16)
            if ctx.from_color and getattr(slots, "single_consumer", False):
                d_hbc = Arena.get(tg + ".d_hbc", (rows, 256), torch.float16, dev)
                slots.d_hb_compact = (d_hbc, d_hbar.data_ptr(), gscale)
            call("spf_head_bwd_tc", C.byref(s), ptr(slots.list), ptr(slots.count), n, ptr(d_rgb.contiguous()), ptr(rgb),
                 ptr(a1), ptr(a2), None if d_hbc is not None else ptr(d_hbar), ptr(dzf), ptr(dz1), ptr(dz2), ptr(dz3),
                 ptr(drb3), ptr(d_hbc), ptr(gscale), stream())
            # every operand is in the tile layout (pe and dz3 with a single k-block)
            (dW4, db4), (dR1f, drb1), (dR1pe, _), (dR2, drb2), (dR3t, _) = _wgrad_multi(
                [(dzf, hb, 256, 256, True), (dz1, f, 256, 256, True), (dz1, pe, 64, 32, False), (dz2, a1, 256, 256, True),
                 (a2, dz3, 64, 16, False)], slots, 1, pool,               # (a2^T @ dz3) = dR3^T, [256,16]
                targets=[(tw[0], tw[1]), (None, tw[3]), (None, None), (tw[4], tw[5]), (None, None)], gscale=gscale)
            dR1 = torch.cat([dR1pe[:, :21], dR1f], dim=1)
            dR3 = dR3t[:, :3].t().contiguous()
            return d_hbar, dW4, db4, dR1, drb1, dR2, drb2, dR3, (None if tw[7] is not None else drb3), None, None, None
        dz3 = Arena.get(tg + ".hdz3", (rows, 4), torch.float32, dev)
        call("spf_head_bwd_f32", C.byref(s), ptr(slots.list), ptr(slots.count), n, ptr(d_rgb.contiguous()), ptr(rgb),
             ptr(a1), ptr(a2), ptr(d_hbar), ptr(dzf), ptr(dz1), ptr(dz2), ptr(dz3), stream())
        # exact mode: the same fp32 FFMA split-K kernel; hbar (by slot) and PE3(dir) (by ray) are read through the list
        dW4, db4 = _wgrad_f32(dzf, hb, 256, slots, 1, idx=slots.list)
        pe = positional_encoding(dirs, 3).contiguous()                                # [R, 21], per ray
        dR1pe, drb1 = _wgrad_f32(dz1, pe, 21, slots, 1, idx=slots.list, idx_div=ctx.Smax)
        dR1f, _ = _wgrad_f32(dz1, f, 256, slots, 1, want_db=False)
        dR1 = torch.cat([dR1pe, dR1f], dim=1)
        dR2, drb2 = _wgrad_f32(dz2, a1, 256, slots, 1)
        dR3, drb3 = _wgrad_f32(dz3, a2, 256, slots, 1, M=3)
        return d_hbar, dW4, db4, dR1, drb1, dR2, drb2, dR3, drb3, None, None, None


class Composite(torch.autograd.Function):
    """Laplace density + alpha compositing + per-ray reductions (density.py:21-30; pointneus_disent.py:701-723,
    765-807, 894-908)."""

    @staticmethod
    def forward(ctx, sdf, rgb_s, beta, delta, t, grad, pidx, nvalid, R, Smax, K, want_normal):
        dev = sdf.device
        ctx.set_materialize_grads(False)   # unused outputs (depth, acc in training) arrive as None, not as zero tensors
        weights = torch.empty(R, Smax, dtype=torch.float32, device=dev)
        rgb = torch.empty(R, 3, dtype=torch.float32, device=dev)
        depth = torch.empty(R, dtype=torch.float32, device=dev)
        acc = torch.empty(R, dtype=torch.float32, device=dev)
        dist = torch.empty(R, dtype=torch.float32, device=dev)
        normal = torch.empty(R, 3, dtype=torch.float32, device=dev) if want_normal else None
        sdf_c, rgb_c, beta_c = sdf.detach().contiguous(), rgb_s.detach().contiguous(), beta.detach().reshape(1).contiguous()
        call("spf_composite_fwd", ptr(sdf_c), ptr(delta), ptr(t), ptr(rgb_c), ptr(grad) if want_normal else None,
             ptr(pidx), K, ptr(nvalid), ptr(beta_c), R, Smax, ptr(weights), ptr(rgb), ptr(depth), ptr(acc), ptr(dist),
             ptr(normal), stream())
        # NB: save a detached alias, never the output object itself (output -> grad_fn -> ctx -> output is a reference
        # cycle through C++ that keeps the whole autograd graph and its AccumulateGrad nodes alive forever)
        ctx.saved_t = (sdf_c, rgb_c, beta_c, delta, t, pidx, nvalid, weights.detach())
        ctx.dims = (R, Smax, K)
        if normal is None:
            normal = torch.zeros(0, 3, device=dev)
        ctx.mark_non_differentiable(normal)
        return weights, rgb, depth, acc, dist, normal

    @staticmethod
    def backward(ctx, d_w, d_rgb, d_depth, d_acc, d_dist, _):
        sdf, rgb_s, beta, delta, t, pidx, nvalid, weights = ctx.saved_t
        R, Smax, K = ctx.dims
        dev = sdf.device
        if d_acc is not None:  # acc = sum_i w_i
            d_w = d_acc[:, None].expand(R, Smax) if d_w is None else d_w + d_acc[:, None]
        c = lambda v: v.contiguous() if v is not None else None
        if d_w is None and d_rgb is None and d_depth is None and d_dist is None:
            return (None,) * 12
        d_sdf = torch.empty(R * Smax, dtype=torch.float32, device=dev)
        d_rgb_s = torch.empty(R * Smax, 3, dtype=torch.float32, device=dev)
        d_beta = torch.zeros(1, dtype=torch.float32, device=dev)
        call("spf_composite_bwd", ptr(sdf), ptr(delta), ptr(t), ptr(rgb_s), ptr(pidx), K, ptr(nvalid), ptr(beta), R,
             Smax, ptr(weights), ptr(c(d_w)), ptr(c(d_rgb)), ptr(c(d_depth)), ptr(c(d_dist)), ptr(d_sdf), ptr(d_rgb_s),
             ptr(d_beta), stream())
        return d_sdf, d_rgb_s, d_beta.reshape(()), None, None, None, None, None, None, None, None, None


class PseudoPointLoss(torch.autograd.Function):
    """pseudo_pts_loss of pointneus_disent.py:765-780 in one piece: expected-depth point of every hit ray -> kNN -> geometry
    field -> masked L1 mean.  Gradients: the geometry latents (scatter-add of the saved Jacobian rows) and ``dist`` (through
    x = cam + dist * dir and d sdf / d x).  Replaces ~40 torch launches of mask / where / sum / sign glue per step."""

    @staticmethod
    def forward(ctx, feat_g, dist, cam_loc, ray_dirs, nvalid, grid, k, r, pack, pts, rbf, owner="", aux=None):
        dev = dist.device
        R = dist.shape[0]
        x = torch.empty(R, 3, dtype=torch.float32, device=dev)
        call("spf_ray_points", ptr(cam_loc), ptr(ray_dirs), ptr(dist.detach().contiguous()), R, ptr(x), stream())
        slots = SlotSet(grid.query_points(x, k, r), "pseudo", owner)
        feat_needs, dist_needs = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        sdf, grad, jw = geo_sdf_raw(pack, slots, x, pts, feat_g.detach(), rbf, dist_needs, feat_needs)
        value = torch.empty(1, dtype=torch.float32, device=dev)
        u_sdf = torch.empty(R, dtype=torch.float32, device=dev)
        u_dist = torch.empty(R, dtype=torch.float32, device=dev) if dist_needs else None
        call("spf_pseudo_loss", ptr(sdf), ptr(grad), ptr(slots.pidx), slots.K, ptr(nvalid), ptr(ray_dirs), R, ptr(value),
             ptr(u_sdf), ptr(u_dist), stream())
        ctx.slots, ctx.jw, ctx.u = slots, jw, (u_sdf, u_dist)
        ctx.feat_shape = feat_g.shape
        ctx.direct = _direct_grad(feat_g)
        if aux is not None:   # the mean's denominator, for the data-parallel global-count normalisation (dist.py)
            aux["pseudo_count"] = ((nvalid > 0) & (slots.pidx[:, 0] >= 0)).sum()
        return value.reshape(())

    @staticmethod
    def backward(ctx, g):
        slots = ctx.slots
        if ctx.jw is not None:
            slots.check_live()
        u_sdf, u_dist = ctx.u
        gfeat = None
        if ctx.jw is not None:
            target = ctx.direct
            if target is None:
                target = gfeat = torch.zeros(ctx.feat_shape, dtype=torch.float32, device=g.device)
            call("spf_sdf_bwd", ptr(slots.list), ptr(slots.count), slots.n, ptr(slots.pidx), slots.K, ptr(ctx.jw),
                 ptr((u_sdf * g).contiguous()), ptr(target), stream())
        d_dist = u_dist * g if u_dist is not None else None
        return gfeat, d_dist, None, None, None, None, None, None, None, None, None, None, None


class TVRegul(torch.autograd.Function):
    """tv_regul (utils.py:221-281) on the cached self-kNN lists of the (static) neural points."""

    @staticmethod
    def forward(ctx, feat_g, pts, self_pidx, first=0, count=None, scale=1.0):
        """`first` / `count` / `scale`: a data-parallel rank evaluates its slice of the points with scale = world size
        (value and gradient), so that the average over the ranks is the full regulariser."""
        N, K = self_pidx.shape
        dev = feat_g.device
        value = torch.empty(1, dtype=torch.float32, device=dev)    # cleared inside spf_tv_fwd_bwd_range
        grad = torch.zeros_like(feat_g, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        call("spf_tv_fwd_bwd_range", ptr(pts), ptr(feat_g.detach().contiguous()), ptr(self_pidx), N, K, int(first),
             int(N if count is None else count), ptr(value), ptr(grad), float(scale), stream())
        ctx.grad = grad
        ctx.direct = _direct_grad(feat_g) if grad is not None else None
        return value.reshape(())

    @staticmethod
    def backward(ctx, g):
        if ctx.direct is not None:
            ctx.direct.addcmul_(ctx.grad, g)   # one pass: grad buffer += unit gradient * upstream scalar
            return None, None, None, None, None, None
        return (ctx.grad * g if ctx.grad is not None else None), None, None, None, None, None


# ------------------------------------------------------------------------------------------------ f4: local loss
def surface_search(sdf, t, cam_loc, ray_dirs, R, Smax, feats=None):
    """spf_local_loss_fwd: first back-facing zero crossing per ray (pointneus_disent.py:586-612) and, with ``feats``
    (see ``local_feature_args``), the per-ray numerator of the feature-consistency loss and its SDF partials."""
    dev = sdf.device
    num = torch.empty(R, dtype=torch.float32, device=dev)
    cross = torch.empty(R, dtype=torch.int32, device=dev)
    d_surface = torch.empty(R, dtype=torch.float32, device=dev)
    g0 = torch.empty(R, dtype=torch.float32, device=dev)
    g1 = torch.empty(R, dtype=torch.float32, device=dev)
    sdf, t, cam_loc, ray_dirs = (v.detach().float().contiguous() for v in (sdf, t, cam_loc, ray_dirs))
    if feats is None:
        fa = (None, None, 0, 0, 0, 32, None, None, 0, 0, 0, None, None)
    else:
        fa = feats["args"]
    call("spf_local_loss_fwd", ptr(sdf), ptr(t), ptr(cam_loc), ptr(ray_dirs), R, Smax, *fa, ptr(num), ptr(cross),
         ptr(d_surface), ptr(g0), ptr(g1), stream())
    return num, cross, d_surface, g0, g1


def local_feature_args(local_data: dict, device) -> dict:
    """Device-side view of the reference's ``local_data`` dict (spurfies/datasets/dtu.py:277-291) for the kernel:
    feat [C,H,W], feat_src [m,C,H,W] (NCHW as the reference holds them, or channels-last storage -- whatever the strides
    say), cam [2,4,4], src_cams [m,2,4,4], size (scalar), center [3].  No copies when the tensors already live on the
    device as fp32; the returned dict keeps them alive."""
    f32 = lambda v: v.to(device=device, dtype=torch.float32)
    feat, src = f32(local_data["feat"]), f32(local_data["feat_src"])
    if src.dim() == 3:
        src = src.unsqueeze(0)
    Cn, H, W = feat.shape
    m = src.shape[0]

    def strides_ok(a, b):
        # both addressed as c * cs + (y * W + x) * ps: NCHW-contiguous or channels-last storage
        return a.stride()[-3:] == b.stride()[-3:] and a.stride(-2) == W * a.stride(-1)
    if not strides_ok(feat, src) or (m > 1 and src.stride(0) <= 0):
        feat, src = feat.contiguous(), src.contiguous()
    dptr = lambda v: v.data_ptr() if v.is_cuda else ptr(v)   # strided (channels-last) views are addressed by stride
    cam = f32(local_data["cam"]).contiguous()
    src_cams = f32(local_data["src_cams"]).reshape(m, 2, 4, 4).contiguous()
    size = f32(torch.as_tensor(local_data["size"])).reshape(-1)[:1].contiguous()
    center = f32(torch.as_tensor(local_data["center"])).reshape(-1)[:3].contiguous()
    keep = (feat, src, cam, src_cams, size, center)
    args = (dptr(feat), dptr(src), int(src.stride(0)) if m > 0 else 0, int(feat.stride(0)), int(feat.stride(2)), int(Cn),
            ptr(cam), ptr(src_cams), int(m), int(H), int(W), ptr(size), ptr(center))
    return {"args": args, "keep": keep, "m": m}


def channels_last_features(local_data: dict, device="cuda") -> dict:
    """Optional one-off re-layout of a view's feature maps to channels-last storage (one 128-byte line per bilinear
    corner instead of 32 sectors); shapes stay [C,H,W] / [m,C,H,W], so the dict is used exactly like the original."""
    out = dict(local_data)
    feat = local_data["feat"].to(device=device, dtype=torch.float32)
    src = local_data["feat_src"].to(device=device, dtype=torch.float32)
    out["feat"] = feat.permute(1, 2, 0).contiguous().permute(2, 0, 1)
    out["feat_src"] = src.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    return out


class LocalLoss(torch.autograd.Function):
    """Feature-consistency loss of the rendered surface (pointneus_disent.py:727-763 + feat_utils.py:377-451):
    mean over (source view, crossing ray) of |1 - cos| gated by in-image and < 0.5; gradient to the dense SDF only."""

    @staticmethod
    def forward(ctx, sdf, t, cam_loc, ray_dirs, feats, R, Smax):
        sdf_c = sdf.detach().contiguous()
        num, cross, d_surface, g0, g1 = surface_search(sdf_c, t, cam_loc, ray_dirs, R, Smax, feats)
        hits = (cross >= 0).sum()
        inv = 1.0 / (feats["m"] * hits.clamp(min=1)).to(torch.float32)
        ctx.saved_t = (cross, g0, g1, inv)
        ctx.dims = (R, Smax)
        ctx.mark_non_differentiable(d_surface, cross)
        return num.sum() * inv, d_surface, cross

    @staticmethod
    def backward(ctx, g, _d, _c):
        cross, g0, g1, inv = ctx.saved_t
        R, Smax = ctx.dims
        d_sdf = torch.empty(R * Smax, dtype=torch.float32, device=cross.device)
        scale = (g * inv).reshape(1).contiguous()
        call("spf_local_loss_bwd", ptr(cross), ptr(g0), ptr(g1), ptr(scale), R, Smax, ptr(d_sdf), stream())
        return d_sdf, None, None, None, None, None, None
