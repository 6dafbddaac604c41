"""oracle/knn.py -- TEST INFRASTRUCTURE ONLY (CPU checker; never imported by the product).

CPU restatement of the reference voxel-grid kNN, ``torch_knnquery`` @ 947957e:

* grid geometry      -> torch_knnquery/torch_knnquery/knnquery.py:32-36, 66-88
* mask / slots / kNN -> C restatement in ``knn_oracle.c`` (cites knnquery.cu lines)
* ragged return      -> knnquery.py:208-285
* brute-force oracle -> torch_knnquery/test/test_queries.py:22-74 (``cdist`` + ``topk``)

``RefVoxelGrid`` mirrors the reference ``VoxelGrid`` API on CPU tensors so the reference's
own Python model code can be driven through it when generating golden vectors
(tests/golden/make_golden.py).

Pinning status: per-sample neighbour sets are pinned against the reference's only KAT
(test_queries.py seed-1234 inputs, brute-force cdist/topk) in tests/test_oracle_knn.py,
and against the unmodified reference CUDA extension (oracle/_ref) on the GPU box in
tests/test_gpu_reference_ext.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile knn_oracle.c with gcc (building the checker is not using it)."""
    so = os.path.join(_HERE, "_build", "libknn_oracle.so")
    src = os.path.join(_HERE, "knn_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", src, "-o", so, "-lm"]
        )
    return so


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        vp, ip, fp = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p
        L.spf_oracle_grid_create.restype = vp
        L.spf_oracle_grid_create.argtypes = [fp, ctypes.c_int, fp, fp, ip, ip]
        L.spf_oracle_grid_destroy.argtypes = [vp]
        L.spf_oracle_grid_stats.argtypes = [vp, ip]
        L.spf_oracle_mask.argtypes = [vp, fp, ctypes.c_int64, ip]
        L.spf_oracle_query.restype = ctypes.c_int64
        L.spf_oracle_query.argtypes = [vp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_float, ip, fp, ip, fp, ip, ip]
        L.spf_oracle_brute.argtypes = [fp, ctypes.c_int, fp, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ip]
        _LIB = L
    return _LIB


def _ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def grid_geometry(points: torch.Tensor, voxel_size, voxel_scale, kernel_size, ranges):
    """Grid origin / cell size / dims exactly as knnquery.py:32-36, 66-88 (torch fp32 ops).

    points: [B,N,3] or [N,3] fp32 (any device). Returns (shift f32[3], scaled_vsize f32[3],
    scaled_vdim i32[3]) as CPU tensors.
    """
    pts = points.detach().reshape(-1, 3).to(torch.float32)
    dev = pts.device
    vscale = torch.tensor(voxel_scale, dtype=torch.float32, device=dev)
    vsize = torch.tensor(voxel_size, dtype=torch.float32, device=dev)
    scaled_vsize = vscale * vsize
    ks = torch.tensor(kernel_size, dtype=torch.int32, device=dev)
    min_xyz, max_xyz = torch.min(pts, dim=0)[0], torch.max(pts, dim=0)[0]
    max_xyz = max_xyz + 0.001
    min_xyz = min_xyz - 0.001
    if ranges is not None:
        rng = torch.tensor(ranges, dtype=torch.float32, device=dev)
        min_xyz = torch.max(torch.stack([min_xyz, rng[:3]], dim=0), dim=0)[0]
        max_xyz = torch.min(torch.stack([max_xyz, rng[3:]], dim=0), dim=0)[0]
    min_xyz = min_xyz - scaled_vsize * ks / 2
    max_xyz = max_xyz + scaled_vsize * ks / 2
    vdim_np = (max_xyz - min_xyz) / vsize
    scaled_vdim = torch.ceil(vdim_np / vscale).type(torch.int32)
    return min_xyz.float().cpu(), scaled_vsize.cpu(), scaled_vdim.cpu()


class OracleGrid:
    """Handle on the C grid (CSR of points per reference-geometry voxel + dilated occupancy)."""

    def __init__(self, points: torch.Tensor, voxel_size, voxel_scale, kernel_size, ranges):
        self.pts = np.ascontiguousarray(points.detach().reshape(-1, 3).cpu().numpy().astype(np.float32))
        shift, vs, dim = grid_geometry(points, voxel_size, voxel_scale, kernel_size, ranges)
        self.shift = np.ascontiguousarray(shift.numpy().astype(np.float32))
        self.vsize = np.ascontiguousarray(vs.numpy().astype(np.float32))
        self.dim = np.ascontiguousarray(dim.numpy().astype(np.int32))
        self.ks = np.ascontiguousarray(np.asarray(kernel_size, dtype=np.int32))
        self.voxel_size = tuple(voxel_size)
        self._h = _lib().spf_oracle_grid_create(_ptr(self.pts), int(self.pts.shape[0]), _ptr(self.shift),
                                                _ptr(self.vsize), _ptr(self.dim), _ptr(self.ks))

    def __del__(self):
        try:
            if self._h:
                _lib().spf_oracle_grid_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def stats(self):
        out = np.zeros(3, dtype=np.int32)
        _lib().spf_oracle_grid_stats(self._h, _ptr(out))
        return {"occupied_voxels": int(out[0]), "max_points_per_voxel": int(out[1]), "points_in_grid": int(out[2])}

    def radius2(self, radius_limit_scale: float) -> float:
        # knnquery.py:247, 258: python double, squared, passed as a C float
        r = radius_limit_scale * max(self.voxel_size[0], self.voxel_size[1])
        return float(np.float32(r ** 2))

    def mask(self, q: torch.Tensor) -> torch.Tensor:
        qn = np.ascontiguousarray(q.detach().reshape(-1, 3).cpu().numpy().astype(np.float32))
        m = np.zeros(qn.shape[0], dtype=np.int32)
        _lib().spf_oracle_mask(self._h, _ptr(qn), qn.shape[0], _ptr(m))
        return torch.from_numpy(m).reshape(q.shape[:-1])

    def query_dense(self, raypos: torch.Tensor, k: int, radius_limit_scale: float, smax: int):
        """raypos [R,D,3] -> dict of dense outputs over all R rays (see knn_oracle.c)."""
        rp = np.ascontiguousarray(raypos.detach().cpu().numpy().astype(np.float32))
        R, D = rp.shape[0], rp.shape[1]
        slot_sample = np.empty((R, smax), dtype=np.int32)
        loc = np.empty((R, smax, 3), dtype=np.float32)
        pidx = np.empty((R, smax, k), dtype=np.int32)
        pd2 = np.empty((R, smax, k), dtype=np.float32)
        m1 = np.empty(R, dtype=np.int8)
        m2 = np.empty(R, dtype=np.int8)
        scanned = _lib().spf_oracle_query(self._h, _ptr(rp), R, D, smax, k,
                                          ctypes.c_float(self.radius2(radius_limit_scale)),
                                          _ptr(slot_sample), _ptr(loc), _ptr(pidx), _ptr(pd2), _ptr(m1), _ptr(m2))
        return {
            "slot_sample": torch.from_numpy(slot_sample), "sample_loc": torch.from_numpy(loc),
            "pidx": torch.from_numpy(pidx), "d2": torch.from_numpy(pd2),
            "ray_mask1": torch.from_numpy(m1), "ray_mask2": torch.from_numpy(m2),
            "candidates_scanned": int(scanned),
        }

    def query_dense_cdist(self, raypos: torch.Tensor, k: int, radius_limit_scale: float, smax: int, chunk: int = 4096):
        """Same contract as query_dense, but the neighbour search is the reference test's brute force
        (test_queries.py:36-41: torch.cdist + topk + radius mask, exact non-matmul distances) in `chunk`-query
        blocks.  This is the "reference CPU path" of BASELINE.md section 2: pure-torch, all host threads."""
        R, D = raypos.shape[0], raypos.shape[1]
        m = self.mask(raypos).bool()                                     # knnquery.cu:171-196
        cum = torch.cumsum(m.int(), -1)
        take = m & (cum <= smax)                                         # knnquery.py:230-231
        slot_sample = torch.full((R, smax), -1, dtype=torch.int32)
        rr, dd = torch.nonzero(take, as_tuple=True)
        ss = (cum[rr, dd] - 1).long()
        slot_sample[rr, ss] = dd.int()
        loc = torch.zeros(R, smax, 3)
        loc[rr, ss] = raypos[rr, dd]
        pidx = torch.full((R, smax, k), -1, dtype=torch.int32)
        pts = torch.from_numpy(self.pts)
        q = raypos[rr, dd]
        r = float(self.radius2(radius_limit_scale)) ** 0.5
        res = []
        for i in range(0, q.shape[0], chunk):
            dist = torch.cdist(q[i:i + chunk][None], pts[None], compute_mode="donot_use_mm_for_euclid_dist")[0]
            top = torch.topk(dist, min(k, dist.shape[-1]), dim=-1, largest=False, sorted=True)
            res.append(torch.where(top.values <= r, top.indices, torch.full_like(top.indices, -1)).int())
        if res:
            pidx[rr, ss] = torch.cat(res, 0)
        m2 = (pidx >= 0).any(-1).any(-1)
        return {"slot_sample": slot_sample, "sample_loc": loc, "pidx": pidx, "ray_mask1": m.any(-1).to(torch.int8),
                "ray_mask2": m2.to(torch.int8), "candidates_scanned": int(q.shape[0]) * int(pts.shape[0])}

    def brute(self, q: torch.Tensor, k: int, radius_limit_scale: float) -> torch.Tensor:
        qn = np.ascontiguousarray(q.detach().reshape(-1, 3).cpu().numpy().astype(np.float32))
        out = np.empty((qn.shape[0], k), dtype=np.int32)
        _lib().spf_oracle_brute(_ptr(self.pts), self.pts.shape[0], _ptr(qn), qn.shape[0], k,
                                ctypes.c_float(self.radius2(radius_limit_scale)), _ptr(out))
        return torch.from_numpy(out)


class RefVoxelGrid(torch.nn.Module):
    """CPU stand-in with the reference ``VoxelGrid`` signature (knnquery.py:21-49, 52-164, 168-285)."""

    def __init__(self, voxel_size, voxel_scale, kernel_size, max_points_per_voxel, max_occ_voxels_per_example,
                 ranges=None):
        super().__init__()
        self.voxel_size, self.voxel_scale, self.kernel_size = voxel_size, voxel_scale, kernel_size
        self.P, self.max_o, self.ranges_t = max_points_per_voxel, max_occ_voxels_per_example, ranges
        self.grid: Optional[OracleGrid] = None
        self._key = None

    def set_pointset(self, points: torch.Tensor, actual_num_points_per_example: torch.Tensor):
        assert points.shape[0] == 1, "B=1 only (the reference itself is broken for B>1: knnquery.cu:271-272)"
        n = int(actual_num_points_per_example.reshape(-1)[0])
        key = (points.data_ptr(), points._version, n)
        if key != self._key:
            # reference derives the bbox from ALL rows (knnquery.py:66) but inserts only the first n
            full = OracleGrid(points, self.voxel_size, self.voxel_scale, self.kernel_size, self.ranges_t)
            if n != points.shape[1]:
                g = OracleGrid.__new__(OracleGrid)
                g.pts = np.ascontiguousarray(full.pts[:n])
                g.shift, g.vsize, g.dim, g.ks, g.voxel_size = full.shift, full.vsize, full.dim, full.ks, full.voxel_size
                g._h = _lib().spf_oracle_grid_create(_ptr(g.pts), n, _ptr(g.shift), _ptr(g.vsize), _ptr(g.dim), _ptr(g.ks))
                full = g
            self.grid, self._key = full, key
            self.caps_bind = None

    def caps_would_bind(self) -> bool:
        st = self.grid.stats()
        return st["max_points_per_voxel"] > self.P or st["occupied_voxels"] > self.max_o

    def query(self, raypos: torch.Tensor, k: int, radius_limit_scale: float, max_shading_points_per_ray: int = 24):
        assert k <= 20, "k cannot be greater than 20"  # knnquery.py:184
        B, R, D = raypos.shape[0], raypos.shape[1], raypos.shape[2]
        assert B == 1
        out = self.grid.query_dense(raypos[0], k, radius_limit_scale, max_shading_points_per_ray)
        keep = out["ray_mask2"].bool()  # knnquery.py:272-280 (rays with >=1 neighbour; subset of ray_mask1)
        return (out["pidx"][keep].to(torch.int32), out["sample_loc"][keep].to(raypos.dtype),
                keep.view(B, R).to(torch.int8))


def brute_force_neighbor_sets(x: torch.Tensor, kp_pos: torch.Tensor, k: int, r: float) -> torch.Tensor:
    """Per-sample neighbour sets the way the reference's own test defines them
    (torch_knnquery/test/test_queries.py:36-41: ``cdist`` -> ``topk(k, largest=False)`` -> keep ``dist < r``).

    x [..., 3], kp_pos [N,3] -> int64 [..., k], ascending point id, -1 padded at the FRONT after sorting
    (compare against ``sorted`` kernel output, as test_queries.py:132-152 does).
    SURVEY D10: the reference test's *slot* rule differs from its kernel's, so only these
    per-sample sets are used as the known answer; slot layout follows knnquery.py:208-231.
    """
    q = x.reshape(-1, 3).to(torch.float32)
    dist = torch.cdist(q[None], kp_pos.reshape(1, -1, 3).to(torch.float32),
                       compute_mode="donot_use_mm_for_euclid_dist")[0]
    top = torch.topk(dist, min(k, dist.shape[-1]), dim=-1, largest=False)
    idx = torch.where(top.values < r, top.indices, torch.full_like(top.indices, -1))
    if idx.shape[-1] < k:
        idx = torch.cat([idx, idx.new_full((idx.shape[0], k - idx.shape[-1]), -1)], -1)
    return idx.sort(dim=-1).values.reshape(*x.shape[:-1], k)
