"""Drop-in for ``torch_knnquery.VoxelGrid`` (torch_knnquery/torch_knnquery/knnquery.py:11-285) on the sm_100a kernels.

Same constructor, ``set_pointset`` and ``query`` signatures and the same ragged return contract
(``sample_pidx`` i32 [R_valid, Smax, k] with -1 padding, ``sample_loc`` f32 [R_valid, Smax, 3] with 0 padding,
``ray_mask`` int8 [B, R] with exactly R_valid ones).  Differences, all deliberate:

* results are deterministic: the K nearest by (d^2, point id), emitted sorted (the reference emits an
  insertion-order-dependent unsorted set);
* ``max_points_per_voxel`` / ``max_occ_voxels_per_example`` never drop points (the reference replaces
  overflow by time-seeded curand reservoir sampling, knnquery.cu:69-79, 158-165); ``caps_exceeded()`` tells
  whether the reference's answer would have been random for this point set;
* the grid is cached on (data_ptr, version, n): the reference rebuilds it 3x per training step from identical
  points (pointneus_disent.py:627, 353, 427);
* B must be 1 (the reference is itself broken for B > 1, knnquery.cu:271-272; Spurfies always uses B = 1);
* ``query_dense`` exposes the dense, sync-free layout the model path uses internally.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import SpfGrid, call, ptr, stream


class VoxelGrid(torch.nn.Module):
    def __init__(self, voxel_size: Tuple[float], voxel_scale: Tuple[float], kernel_size: Tuple[int],
                 max_points_per_voxel: int, max_occ_voxels_per_example: float, ranges: Optional[Tuple[int]] = None):
        super().__init__()
        self.vsize_tup = tuple(voxel_size)
        self.register_buffer("vscale", torch.tensor(voxel_scale, dtype=torch.float32), persistent=False)
        self.register_buffer("vsize", torch.tensor(voxel_size, dtype=torch.float32), persistent=False)
        self.register_buffer("scaled_vsize", self.vscale * self.vsize, persistent=False)
        self.register_buffer("kernel_size", torch.tensor(kernel_size, dtype=torch.int32), persistent=False)
        self.P = max_points_per_voxel
        self.max_o = max_occ_voxels_per_example
        self.register_buffer("ranges_original",
                             torch.tensor(ranges, dtype=torch.float32) if ranges is not None else None,
                             persistent=False)
        self._ks = tuple(int(k) for k in kernel_size)
        self._key = None
        self._grid: Optional[SpfGrid] = None
        self._stats = None
        self._search_radius = None
        self.use_search_grid = True   # False: always scan the 27 reference voxels (same results, more candidates)

    # ------------------------------------------------------------------ set_pointset (knnquery.py:52-164)
    def set_pointset(self, points: torch.Tensor, actual_num_points_per_example: torch.Tensor) -> None:
        assert points.is_cuda
        if points.dim() != 3 or points.shape[0] != 1:
            raise _lib.SpfError("VoxelGrid.set_pointset: expected points of shape [1, N, 3] (B = 1)")
        if points.dtype != torch.float32 or not points.is_contiguous():
            raise _lib.SpfError("points must be a contiguous float32 tensor")
        n_t = actual_num_points_per_example
        key = (points.data_ptr(), points._version, tuple(points.shape),
               n_t.data_ptr() if torch.is_tensor(n_t) else int(n_t), n_t._version if torch.is_tensor(n_t) else 0)
        if key == self._key and self._grid is not None:
            return
        n = int(n_t.reshape(-1)[0]) if torch.is_tensor(n_t) else int(n_t)
        dev = points.device
        self.points = points
        self.B, self.N = 1, points.shape[1]
        # grid geometry: identical torch fp32 ops to knnquery.py:66-88 (run once per point set)
        flat = points.flatten(0, 1)
        min_xyz, max_xyz = torch.min(flat, dim=0)[0], torch.max(flat, dim=0)[0]
        max_xyz = max_xyz + 0.001
        min_xyz = min_xyz - 0.001
        if self.ranges_original is not None:
            ro = self.ranges_original.to(dev)
            min_xyz = torch.max(torch.stack([min_xyz, ro[:3]], dim=0), dim=0)[0]
            max_xyz = torch.min(torch.stack([max_xyz, ro[3:]], dim=0), dim=0)[0]
        svs, ks = self.scaled_vsize.to(dev), self.kernel_size.to(dev)
        min_xyz = min_xyz - svs * ks / 2
        max_xyz = max_xyz + svs * ks / 2
        self.ranges = torch.cat([min_xyz, max_xyz], dim=-1).float()
        vdim = (max_xyz - min_xyz) / self.vsize.to(dev)
        self.scaled_vdim = torch.ceil(vdim / self.vscale.to(dev)).type(torch.int32)
        host = torch.cat([self.ranges[:3], svs, self.scaled_vdim.float()]).cpu()  # one D2H of 9 floats per build
        dim = [int(v) for v in host[6:9]]
        g = SpfGrid()
        for a in range(3):
            g.shift[a] = float(host[a])
            g.vsize[a] = float(host[3 + a])
            g.dim[a] = dim[a]
            g.ks[a] = self._ks[a]
        G = dim[0] * dim[1] * dim[2]
        if G <= 0 or G >= 2 ** 31:
            raise _lib.SpfError(f"degenerate voxel grid {dim}")
        g.n_points, g.n_cells = n, G
        self._cell_start = torch.empty(G + 1, dtype=torch.int32, device=dev)
        self._sorted = torch.empty(max(n, 1), 4, dtype=torch.float32, device=dev)
        self._hit = torch.empty(G, dtype=torch.uint8, device=dev)
        self._stats_dev = torch.zeros(4, dtype=torch.int32, device=dev)
        ws = torch.empty(_lib.lib.spf_grid_workspace_bytes(n, G), dtype=torch.uint8, device=dev)
        g.cell_start, g.sorted, g.hit = self._cell_start.data_ptr(), self._sorted.data_ptr(), self._hit.data_ptr()
        call("spf_grid_build", C.byref(g), ptr(points), ptr(self._cell_start), ptr(self._sorted), ptr(self._hit),
             ptr(self._stats_dev), ptr(ws), ws.numel(), stream())
        self._grid, self._key, self._stats, self._search_radius = g, key, None, None
        # kernel-family hint for the ray-slot kNN (spf_grid.dense_cloud): one more small D2H per point-set build
        g.dense_cloud = int(self.stats()["max_points_per_voxel"] > 256)
        self.grid_dim = tuple(dim)
        self.d_coord_shift = self.ranges[:3]

    def stats(self) -> dict:
        if self._stats is None:
            s = self._stats_dev.cpu()
            self._stats = {"occupied_voxels": int(s[0]), "max_points_per_voxel": int(s[1]), "points_in_grid": int(s[2])}
        return self._stats

    def caps_exceeded(self) -> bool:
        """True when the reference's P / max_o caps would bind, i.e. when ITS result is a random subset."""
        s = self.stats()
        return s["max_points_per_voxel"] > self.P or s["occupied_voxels"] > self.max_o

    def radius2(self, radius_limit_scale: float) -> float:
        radius_limit = radius_limit_scale * max(self.vsize_tup[0], self.vsize_tup[1])  # knnquery.py:247
        self._ensure_search_grid(radius_limit)
        return radius_limit ** 2

    def _ensure_search_grid(self, radius: float) -> None:
        """Build (once per point set and radius) the SEARCH grid: cubic cells a hair larger than the query radius, so a
        query scans the 27 cells covering its radius ball instead of the 27 (2.25x wider) reference voxels.  Used by the
        kernels only when it cannot change the result (radius <= reference voxel edge, see spf_grid in the header)."""
        g = self._grid
        if g is None or not self.use_search_grid or radius <= 0.0 or radius > min(g.vsize) or self._search_radius == radius:
            return
        cell = float(radius) * 1.001
        dims = [max(1, int(math.ceil(g.dim[a] * g.vsize[a] / cell))) for a in range(3)]
        G = dims[0] * dims[1] * dims[2]
        if G >= 2 ** 31:
            return
        dev = self.points.device
        self._search_cell_start = torch.empty(G + 1, dtype=torch.int32, device=dev)
        self._search_sorted = torch.empty(max(g.n_points, 1), 4, dtype=torch.float32, device=dev)
        g.search_cell = cell
        for a in range(3):
            g.search_dim[a] = dims[a]
        ws = torch.empty(_lib.lib.spf_grid_workspace_bytes(g.n_points, G), dtype=torch.uint8, device=dev)
        call("spf_grid_build_search", C.byref(g), ptr(self.points), ptr(self._search_cell_start), ptr(self._search_sorted),
             ptr(ws), ws.numel(), stream())
        g.search_cell_start, g.search_sorted = self._search_cell_start.data_ptr(), self._search_sorted.data_ptr()
        self._search_radius = radius

    @property
    def handle(self) -> SpfGrid:
        if self._grid is None:
            raise _lib.SpfError("VoxelGrid: set_pointset has not been called")
        return self._grid

    # ------------------------------------------------------------------ dense, sync-free query
    def query_dense(self, raypos: torch.Tensor, k: int, radius_limit_scale: float, max_shading_points_per_ray: int):
        """raypos [R, D, 3] -> (pidx i32 [R,Smax,k], sample_loc f32 [R,Smax,3], slot_sample i32 [R,Smax],
        ray_nvalid i32 [R] = slots with >= 1 neighbour).  No host synchronisation."""
        assert k <= 20, "k cannot be greater than 20"  # knnquery.py:184
        if raypos.dtype != torch.float32:
            raise _lib.SpfError("raypos must be float32")
        R, D = raypos.shape[0], raypos.shape[1]
        S = int(max_shading_points_per_ray)
        dev = raypos.device
        slot_sample = torch.empty(R, S, dtype=torch.int32, device=dev)
        loc = torch.empty(R, S, 3, dtype=torch.float32, device=dev)
        n_slots = torch.empty(R, dtype=torch.int32, device=dev)
        pidx = torch.empty(R, S, k, dtype=torch.int32, device=dev)
        nvalid = torch.empty(R, dtype=torch.int32, device=dev)
        g = C.byref(self.handle)
        call("spf_mask_slots", g, ptr(raypos), R, D, S, ptr(slot_sample), ptr(loc), ptr(n_slots), stream())
        call("spf_knn_slots", g, ptr(loc), ptr(n_slots), R, S, k, self.radius2(radius_limit_scale), ptr(pidx),
             ptr(nvalid), stream())
        return pidx, loc, slot_sample, nvalid

    def query_points(self, q: torch.Tensor, k: int, radius_limit_scale: float, skip: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Point queries (D = 1, Smax = 1 semantics): q [Q,3] -> pidx i32 [Q,k] (-1 rows where masked out / empty).
        ``skip`` (int32 [1] on the device, optional): when non-zero at launch time nothing is searched and every row is
        -1 (spf_knn_points_pred)."""
        assert k <= 20, "k cannot be greater than 20"
        if q.dtype != torch.float32:
            raise _lib.SpfError("q must be float32")
        Q = q.shape[0]
        pidx = torch.empty(Q, k, dtype=torch.int32, device=q.device)
        if skip is not None:
            assert skip.dtype == torch.int32 and skip.is_cuda and skip.numel() == 1
        call("spf_knn_points_pred", C.byref(self.handle), ptr(q), Q, k, self.radius2(radius_limit_scale), ptr(pidx),
             ptr(skip), stream())
        return pidx

    def mask_points(self, q: torch.Tensor) -> torch.Tensor:
        m = torch.empty(q.shape[:-1], dtype=torch.int32, device=q.device)
        call("spf_mask_points", C.byref(self.handle), ptr(q), m.numel(), ptr(m), stream())
        return m

    # ------------------------------------------------------------------ query (knnquery.py:168-285)
    def query(self, raypos: torch.Tensor, k: int, radius_limit_scale: float,
              max_shading_points_per_ray: Optional[int] = 24):
        assert k <= 20, "k cannot be greater than 20"
        if raypos.dim() != 4 or raypos.shape[0] != 1:
            raise _lib.SpfError("VoxelGrid.query: expected raypos of shape [1, R, D, 3] (B = 1)")
        if not raypos.is_cuda or not raypos.is_contiguous():
            raise _lib.SpfError("raypos must be a contiguous CUDA tensor")
        R = raypos.shape[1]
        pidx, loc, _, nvalid = self.query_dense(raypos[0], k, radius_limit_scale, max_shading_points_per_ray)
        ray_mask = nvalid > 0  # knnquery.py:272-280 (a ray without any neighbour is dropped)
        # ragged contract of the reference: compaction (one host sync, only at this API boundary)
        return pidx[ray_mask], loc[ray_mask], ray_mask.view(1, R).to(torch.int8)
