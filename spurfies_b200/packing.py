"""Weight images for the tcgen05 kernels (mlp_tc.cu).

A B operand of ``tcgen05.mma`` is an [N x K] matrix with K contiguous ("K-major"), i.e. exactly a torch
``nn.Linear.weight`` [out, in] for a forward layer, or its transpose for a dgrad layer.  The kernels load it with
plain bulk copies (``cp.async.bulk``), so the image in global memory must already be in the shared-memory layout
the MMA descriptor expects: k-blocks of 64 bf16 columns; inside a k-block one 128-byte row per n; the eight
16-byte chunks of a row XOR-swizzled by (n & 7) (the hardware's 128B swizzle).  Byte size: ceil(K/64) * N * 128.
"""
from __future__ import annotations

import torch


PACK_F16 = 2   # SPF_PACK_F16: forward weight images are fp16, gradient-chain images bf16 (csrc/umma.cuh)


def pack_sw128(W: torch.Tensor, n_pad: int | None = None, dtype=torch.bfloat16) -> torch.Tensor:
    """W [N, K] (any float dtype, any device) -> uint8 image [ceil(K/64) * n_pad * 128] on the same device."""
    N, K = W.shape
    n_pad = n_pad or N
    assert n_pad % 8 == 0 and n_pad >= N
    nkb = (K + 63) // 64
    Wb = torch.zeros(n_pad, nkb * 64, dtype=dtype, device=W.device)
    Wb[:N, :K] = W.detach().float().clamp(-65504.0, 65504.0).to(dtype)
    t = Wb.view(n_pad, nkb, 8, 8).permute(1, 0, 2, 3)                      # [kb, n, chunk, 8]
    n = torch.arange(n_pad, device=W.device)
    c = torch.arange(8, device=W.device)
    src = (c[None, :] ^ (n[:, None] & 7))                                  # out chunk c holds in chunk c ^ (n & 7)
    t = torch.gather(t, 2, src[None, :, :, None].expand(nkb, n_pad, 8, 8))
    return t.contiguous().view(torch.uint8).reshape(-1)


def pack_sw128_dev(out: torch.Tensor, W: torch.Tensor, N: int, K: int, transpose: bool = False, n_pad: int | None = None,
                   col_off: int = 0, row_off_bytes: int = 0, batch: list | None = None, f16: bool = False) -> None:
    """Kernel version (spf_pack_sw128) for per-step packing of trainable weights: writes the image of
    W[:, col_off:col_off+K] ([N, K]) -- or, with transpose, of W[:, col_off:col_off+N]^T where W is [K, .] -- into the
    uint8 buffer `out` starting at byte `row_off_bytes`.  W must be fp32, CUDA, with unit column stride.  With `batch`
    (a list) the job is only recorded; ``pack_flush(batch)`` then packs all recorded images in one launch."""
    import ctypes as C
    from . import _lib
    assert W.dtype == torch.float32 and W.is_cuda and W.stride(1) == 1
    n_pad = n_pad or N
    if batch is not None:   # collected; one launch for all of them in pack_flush
        batch.append((W.data_ptr() + 4 * col_off, out.data_ptr() + row_off_bytes, int(W.stride(0)), int(N), int(K),
                      int(bool(transpose)) | (PACK_F16 if f16 else 0), int(n_pad), W))
        return
    src = C.c_void_p(W.data_ptr() + 4 * col_off)
    dst = C.c_void_p(out.data_ptr() + row_off_bytes)
    _lib.call("spf_pack_sw128", src, int(W.stride(0)), int(N), int(K), int(bool(transpose)) | (PACK_F16 if f16 else 0), int(n_pad), dst,
              _lib.stream())


def pack_flush(batch: list) -> None:
    """One spf_pack_sw128_batch launch for the jobs collected by ``pack_sw128_dev(..., batch=batch)``."""
    import ctypes as C
    from . import _lib
    for i in range(0, len(batch), 16):
        part = batch[i:i + 16]
        arr = (_lib.PackJob * len(part))()
        for j, (src, dst, ld, N, K, tr, n_pad, _keep) in enumerate(part):
            arr[j].W, arr[j].out, arr[j].ld, arr[j].N, arr[j].K, arr[j].transpose, arr[j].n_pad = src, dst, ld, N, K, tr, n_pad
        _lib.call("spf_pack_sw128_batch", C.cast(arr, C.c_void_p), len(part), _lib.stream())
    batch.clear()


def image_bytes(N_pad: int, K: int) -> int:
    return ((K + 63) // 64) * N_pad * 128
