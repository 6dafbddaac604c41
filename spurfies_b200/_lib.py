"""ctypes binding of libspurfies_b200.so (include/spurfies_b200.h).  No CPU fallback: if the library is
missing and cannot be built, importing this module raises."""
from __future__ import annotations

import ctypes as C
import os

import torch

from .build import LIB, build_library

_here = os.path.dirname(os.path.abspath(__file__))


def _load():
    path = os.environ.get("SPF_LIBRARY")   # development only: an instrumented build of the same sources
    if path:
        return C.CDLL(path)
    path = LIB
    if not os.path.exists(path):
        path = build_library()
    return C.CDLL(path)


lib = _load()


class SpfGrid(C.Structure):
    _fields_ = [("shift", C.c_float * 3), ("vsize", C.c_float * 3), ("dim", C.c_int32 * 3), ("ks", C.c_int32 * 3),
                ("n_points", C.c_int32), ("n_cells", C.c_int32), ("cell_start", C.c_void_p), ("sorted", C.c_void_p),
                ("hit", C.c_void_p), ("search_cell", C.c_float), ("search_dim", C.c_int32 * 3),
                ("search_cell_start", C.c_void_p), ("search_sorted", C.c_void_p), ("dense_cloud", C.c_int32)]


class GeoWeightsF32(C.Structure):
    _fields_ = [("w1t", C.c_void_p), ("b1", C.c_void_p), ("w2t", C.c_void_p), ("b2", C.c_void_p),
                ("w3t", C.c_void_p), ("b3", C.c_void_p), ("w4t", C.c_void_p), ("b4", C.c_void_p),
                ("v5", C.c_void_p), ("c5", C.c_float),
                ("w1", C.c_void_p), ("w2", C.c_void_p), ("w3", C.c_void_p), ("w4", C.c_void_p)]


class GeoWeightsTC(C.Structure):
    _fields_ = [("w1p", C.c_void_p), ("w2p", C.c_void_p), ("w3p", C.c_void_p), ("w4p", C.c_void_p),
                ("w4tp", C.c_void_p), ("w3tp", C.c_void_p), ("w2tp", C.c_void_p), ("w1tp", C.c_void_p),
                ("b1", C.c_void_p), ("b2", C.c_void_p), ("b3", C.c_void_p), ("b4", C.c_void_p), ("v5", C.c_void_p),
                ("c5", C.c_float)]


class ColorWeightsTC(C.Structure):
    _fields_ = [("w1p", C.c_void_p), ("w2p", C.c_void_p), ("w3p", C.c_void_p), ("w3tp", C.c_void_p),
                ("w2tp", C.c_void_p), ("w1ftp", C.c_void_p), ("b1", C.c_void_p), ("b2", C.c_void_p), ("b3", C.c_void_p)]


class HeadWeightsTC(C.Structure):
    _fields_ = [("w4p", C.c_void_p), ("r1fp", C.c_void_p), ("r2p", C.c_void_p), ("r3p", C.c_void_p),
                ("r3tp", C.c_void_p), ("r2tp", C.c_void_p), ("r1ftp", C.c_void_p), ("w4tp", C.c_void_p),
                ("b4", C.c_void_p), ("rb2", C.c_void_p), ("rb3", C.c_void_p)]


class PackJob(C.Structure):
    _fields_ = [("W", C.c_void_p), ("out", C.c_void_p), ("ld", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
                ("transpose", C.c_int32), ("n_pad", C.c_int32), ("reserved", C.c_int32)]


class WgradJob(C.Structure):
    _fields_ = [("dz", C.c_void_p), ("act", C.c_void_p), ("dW", C.c_void_p), ("db", C.c_void_p), ("lda", C.c_int32),
                ("N", C.c_int32), ("fmt", C.c_int32), ("reserved", C.c_int32)]


class ColorWeightsF32(C.Structure):
    _fields_ = [("w1t", C.c_void_p), ("b1", C.c_void_p), ("w2t", C.c_void_p), ("b2", C.c_void_p),
                ("w3t", C.c_void_p), ("b3", C.c_void_p), ("w1", C.c_void_p), ("w2", C.c_void_p), ("w3", C.c_void_p)]


class HeadWeightsF32(C.Structure):
    _fields_ = [("w4t", C.c_void_p), ("b4", C.c_void_p), ("r1t", C.c_void_p), ("rb1", C.c_void_p),
                ("r2t", C.c_void_p), ("rb2", C.c_void_p), ("r3t", C.c_void_p), ("rb3", C.c_void_p),
                ("w4", C.c_void_p), ("r1", C.c_void_p), ("r2", C.c_void_p), ("r3", C.c_void_p)]


lib.spf_version.restype = C.c_char_p
lib.spf_last_cuda_error.restype = C.c_char_p
lib.spf_grid_workspace_bytes.restype = C.c_size_t
lib.spf_grid_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
lib.spf_compact_workspace_bytes.restype = C.c_size_t
lib.spf_compact_workspace_bytes.argtypes = [C.c_int64]
lib.spf_optim_workspace_bytes.restype = C.c_size_t
lib.spf_optim_workspace_bytes.argtypes = []
lib.spf_loss_workspace_bytes.restype = C.c_size_t
lib.spf_loss_workspace_bytes.argtypes = []
lib.spf_voxelize_workspace_bytes.restype = C.c_size_t
lib.spf_voxelize_workspace_bytes.argtypes = [C.c_int32]

_P, _I, _L, _F, _Z, _D = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t, C.c_double
_SIGS = {
    "spf_grid_build": [_P, _P, _P, _P, _P, _P, _P, _Z, _P],
    "spf_grid_build_search": [_P, _P, _P, _P, _P, _Z, _P],
    "spf_mask_slots": [_P, _P, _I, _I, _I, _P, _P, _P, _P],
    "spf_knn_slots": [_P, _P, _P, _I, _I, _I, _F, _P, _P, _P],
    "spf_knn_points": [_P, _P, _L, _I, _F, _P, _P],
    "spf_knn_points_pred": [_P, _P, _L, _I, _F, _P, _P, _P],
    "spf_knn_set_algo": [_I],
    "spf_grad_scale": [_P, _L, _F, _P, _P, _P],
    "spf_mask_points": [_P, _P, _L, _P, _P],
    "spf_compact_valid": [_P, _L, _I, _P, _P, _P, _Z, _P],
    "spf_ray_prep": [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P],
    "spf_sdf_fwd_f32": [_P, _P, _P, _L, _P, _P, _I, _P, _P, _F, _P, _P, _P, _P],
    "spf_sdf_bwd": [_P, _P, _L, _P, _I, _P, _P, _P, _P],
    "spf_color_fwd_f32": [_P, _P, _P, _L, _P, _P, _I, _P, _P, _F, _P, _P, _P, _P, _P, _P, _P],
    "spf_color_bwd_f32": [_P, _P, _P, _L, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "spf_head_fwd_f32": [_P, _P, _P, _L, _P, _P, _I, _P, _P, _P, _P, _P],
    "spf_head_bwd_f32": [_P, _P, _P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "spf_composite_fwd": [_P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    "spf_composite_bwd": [_P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "spf_sampler_coarse": [_P, _P, _F, _F, _P, _P, _I, _I, _P, _P, _P],
    "spf_sampler_iter": [_P, _P, _I, _I, _P, _F, _I, _F, _F, _I, _P, _I, _I, _P, _P, _F, _F, _P, _I, _P, _P, _P, _P,
                         _P, _P],
    "spf_sampler_iter_pred": [_P, _P, _I, _I, _P, _F, _I, _F, _F, _I, _P, _I, _I, _P, _P, _F, _F, _P, _I, _P, _P, _P, _P,
                              _P, _I, _P],
    "spf_wgrad_f32": [_P, _I, _I, _P, _I, _I, _P, _I, _P, _I, _L, _P, _P, _P],
    "spf_sampler_merge": [_P, _P, _I, _P, _P, _I, _I, _P, _P, _P],
    "spf_tv_fwd_bwd": [_P, _P, _P, _I, _I, _P, _P, _F, _P],
    "spf_tv_fwd_bwd_range": [_P, _P, _P, _I, _I, _I, _I, _P, _P, _F, _P],
    "spf_camera_rays": [_P, _P, _P, _I, _P, _P, _P, _P],
    "spf_ray_points": [_P, _P, _P, _I, _P, _P],
    "spf_pseudo_loss": [_P, _P, _P, _I, _P, _P, _I, _P, _P, _P, _P],
    "spf_volsdf_loss": [_P, _P, _P, _P, _I, _P, _P, _L, _I, _I, _P, _P, _P, _F, _F, _F, _F, _F, _P, _P, _P, _P, _Z, _P],
    "spf_grad_sumsq": [_P, _L, _F, _P, _P, _Z, _P],
    "spf_adam_step": [_P, _P, _P, _P, _L, _P, _P, _F, _F, _D, _D, _F, _I, _P, _P],
    "spf_grid_points_mask": [_P, _P, _P, _P, _I, _I, _I, _L, _L, _F, _P, _P, _P, _P, _I, _P],
    "spf_grid_points_mask_cyclic": [_P, _P, _P, _P, _I, _I, _I, _L, _L, _L, _I, _I, _P, _F, _P, _P, _P, _P, _I, _P],
    "spf_scatter_f32": [_P, _P, _I, _P, _P],
    "spf_local_loss_fwd": [_P, _P, _P, _P, _I, _I, _P, _P, _L, _L, _L, _I, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P,
                           _P],
    "spf_local_loss_bwd": [_P, _P, _P, _P, _I, _I, _P, _P],
    "spf_voxelize_closest": [_P, _I, _F, _F, _F, _F, _I, _P, _P, _P, _I, _P, _P, _Z, _P],
    "spf_tc_gemm_test": [_P, _P, _I, _I, _P, _P],
    "spf_sdf_fwd_tc": [_P, _P, _P, _L, _P, _P, _I, _P, _P, _F, _P, _P, _P, _P],
    "spf_wgrad_tc": [_P, _P, _I, _I, _P, _I, _L, _I, _P, _P, _P],
    "spf_wgrad_tc_multi": [_P, _I, _P, _I, _L, _P, _P],
    "spf_head_fwd_tc": [_P, _P, _P, _L, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P],
    "spf_head_bwd_tc": [_P, _P, _P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "spf_pack_sw128": [_P, _I, _I, _I, _I, _I, _P, _P],
    "spf_pack_sw128_batch": [_P, _I, _P],
    "spf_head_zpe": [_P, _P, _I, _P, _I, _P, _P],
    "spf_color_fwd_tc": [_P, _P, _P, _L, _P, _P, _I, _P, _P, _F, _P, _P, _P, _P, _P, _P, _P, _P],
    "spf_color_bwd_tc": [_P, _P, _P, _L, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
}
for _n, _a in _SIGS.items():
    _f = getattr(lib, _n)
    _f.restype = C.c_int
    _f.argtypes = _a

if os.environ.get("SPF_KNN_ALGO"):   # development knob: A/B the kNN kernel families under bench.py (spf_knn_set_algo)
    lib.spf_knn_set_algo(int(os.environ["SPF_KNN_ALGO"]))

EXPORTED = ["spf_version", "spf_last_cuda_error", "spf_grid_workspace_bytes", "spf_compact_workspace_bytes",
            "spf_optim_workspace_bytes", "spf_voxelize_workspace_bytes", "spf_loss_workspace_bytes", *_SIGS]


class SpfError(RuntimeError):
    pass


_CODES = {-1: "SPF_ERR_INVALID", -2: "SPF_ERR_UNSUPPORTED", -3: "SPF_ERR_WORKSPACE", -4: "SPF_ERR_CUDA"}


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL). Tensors must be CUDA + contiguous, like the reference
    extension requires (torch_knnquery/src/knnquery.cu:13-15)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise SpfError("tensor must be a CUDA tensor")
    if not t.is_contiguous():
        raise SpfError("tensor is not contiguous")
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---- optional per-entry-point device timing (CUDA events on the launching stream), used by bench.py
_PROFILE = {"on": False, "events": {}, "calls": {}}


def profile_reset(on: bool):
    _PROFILE["on"], _PROFILE["events"], _PROFILE["calls"] = on, {}, {}


def profile_collect():
    """-> {entry point: {"calls", "ms" (total), "max_ms"}}; synchronises."""
    torch.cuda.synchronize()
    out = {}
    for name, evs in _PROFILE["events"].items():
        times = [a.elapsed_time(b) for a, b in evs]
        out[name] = {"calls": len(times), "ms": sum(times), "max_ms": max(times)}
    return out


# SPF_NVTX=1: an NVTX range per C-ABI call, named after the entry point (ncu --nvtx --nvtx-include "spf_sdf_fwd_tc/" ...)
_NVTX = os.environ.get("SPF_NVTX") == "1"


def call(name, *args):
    if not _NVTX:
        return _call(name, *args)
    torch.cuda.nvtx.range_push(name)
    try:
        return _call(name, *args)
    finally:
        torch.cuda.nvtx.range_pop()


def _call(name, *args):
    if _PROFILE["on"]:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = getattr(lib, name)(*args)
        b.record()
        _PROFILE["events"].setdefault(name, []).append((a, b))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = _CODES.get(rc, str(rc))
        if rc == -4:
            msg += ": " + lib.spf_last_cuda_error().decode()
        raise SpfError(f"{name} failed: {msg}")
