"""TEST / BENCH INFRASTRUCTURE ONLY (never imported by spurfies_b200/): the reference's GPU path on the same B200.

The reference itself cannot travel to the GPU box (its Python lives under /root/reference, which does not exist there)
and pins torch 2.0; what CAN run there is

  * its own CUDA kernels, compiled unmodified from where they lie by oracle/build_ref.sh into
    oracle/_ref/knnquery_cuda*.so (git-ignored, shipped with the snapshot), and
  * oracle/hotpath.py -- the restatement of its torch graph that tests/test_oracle_hotpath.py pins to outputs of the
    reference's own Python at 1e-5 -- which is plain torch and therefore runs on CUDA too (cuBLAS fp32 GEMMs,
    index_add_, cumsum, searchsorted, sort, autograd: the very ops the reference runs on its GPU).

``RefExtVoxelGrid`` drives the compiled reference kernels the way torch_knnquery/torch_knnquery/knnquery.py:52-285
drives them (same call sequence, same allocations per call, same host syncs) behind the same public API
(``set_pointset`` / ``query``).  ``ApiGrid`` adapts ANY object with that API to the oracle's ``query_dense`` contract,
following the reference's glue (spurfies/model/utils.py:90-113 + pointneus_disent.py:627: the point set is re-inserted
before every query).  Plugging ``spurfies_b200.knnquery.VoxelGrid`` into the same adapter is the Level-1 drop-in check:
the reference-side torch graph runs unchanged on top of the product's kNN.

Used by tests/test_gpu_reference_path.py (drop-in equivalence) and bench.py (``reference_gpu``: the reference's training
step and its kNN query timed on the same GPU, next to ours).
"""
from __future__ import annotations

import glob
import importlib.util
import os
import time
from contextlib import contextmanager

import torch

from . import hotpath as H
from .knn import grid_geometry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_reference_ext():
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "knnquery_cuda*.so"))
    if not so:
        return None
    spec = importlib.util.spec_from_file_location("knnquery_cuda", so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class RefExtVoxelGrid:
    """torch_knnquery.VoxelGrid (knnquery.py:21-285) on the unmodified compiled reference kernels, B = 1."""

    def __init__(self, voxel_size, voxel_scale, kernel_size, max_points_per_voxel, max_occ_voxels_per_example, ranges=None):
        self.ext = load_reference_ext()
        if self.ext is None:
            raise RuntimeError("oracle/_ref/knnquery_cuda*.so is not built (oracle/build_ref.sh needs /root/reference)")
        self.voxel_size, self.voxel_scale, self.kernel = tuple(voxel_size), tuple(voxel_scale), tuple(kernel_size)
        self.P, self.max_o, self.ranges = int(max_points_per_voxel), int(max_occ_voxels_per_example), ranges

    def set_pointset(self, points: torch.Tensor, actual_num_points_per_example: torch.Tensor) -> None:
        """knnquery.py:52-164: geometry from the points, five fresh tables, three kernels."""
        assert points.is_cuda and points.shape[0] == 1
        dev = points.device
        self.points = points
        B, N = points.shape[0], points.shape[1]
        shift, svs, vdim = grid_geometry(points, self.voxel_size, self.voxel_scale, self.kernel, self.ranges)
        self.shift, self.svs, self.vdim = shift.to(dev), svs.to(dev), vdim.to(dev)
        dims = [int(v) for v in vdim]                                         # host sync, as knnquery.py:84-90
        self.G = dims[0] * dims[1] * dims[2]
        self.ks = torch.tensor(self.kernel, dtype=torch.int32, device=dev)
        self.coor_occ = torch.zeros([B] + dims, dtype=torch.int32, device=dev)
        self.occ_2_pnts = torch.full([B, self.max_o, self.P], -1, dtype=torch.int32, device=dev)
        occ_2_coor = torch.full([B, self.max_o, 3], -1, dtype=torch.int32, device=dev)
        self.occ_numpnts = torch.zeros([B, self.max_o], dtype=torch.int32, device=dev)
        coor_2_occ = torch.full([B] + dims, -1, dtype=torch.int32, device=dev)
        occ_idx = torch.zeros([B], dtype=torch.int32, device=dev)
        n = actual_num_points_per_example
        sec = int(round(time.time() * 1000))
        self.ext.find_occupied_voxels(points, n, B, N, self.shift, self.svs, self.vdim, self.G, self.max_o, occ_idx,
                                      coor_2_occ, occ_2_coor, sec)
        self.coor_2_occ = torch.full([B] + dims, -1, dtype=torch.int32, device=dev)
        self.ext.create_coor_occ_maps(B, self.vdim, self.ks, self.G, self.max_o, occ_idx, self.coor_occ, self.coor_2_occ,
                                      occ_2_coor)
        sec = int(round(time.time() * 1000))
        self.ext.assign_points_to_occ_voxels(points, n, B, N, self.P, self.shift, self.svs, self.vdim, self.G, self.max_o,
                                             self.coor_2_occ, self.occ_2_pnts, self.occ_numpnts, sec)
        self.occ_idx = occ_idx

    def query(self, raypos: torch.Tensor, k: int, radius_limit_scale: float, max_shading_points_per_ray: int = 24):
        """knnquery.py:168-285: mask, first-Smax slots, kNN; rows of rays without any neighbour are dropped."""
        dev = raypos.device
        B, R, D = 1, raypos.size(1), raypos.size(2)
        S = max_shading_points_per_ray
        assert k <= 20
        mask = torch.zeros([B, R, D], dtype=torch.int32, device=dev)
        self.ext.create_raypos_mask(raypos, self.coor_occ, B, R, D, self.G, self.shift, self.vdim, self.svs, mask)
        mask = mask.view(R, D)
        r2b = torch.zeros(R, dtype=torch.int32, device=dev)
        rp = raypos.view(R, D, 3)
        ray_mask_1 = mask.max(-1)[0] > 0
        R_valid = int(ray_mask_1.sum())                                              # host sync (knnquery.py:214)
        loc = torch.zeros([R_valid, S, 3], dtype=raypos.dtype, device=dev)
        pidx = torch.full([R_valid, S, k], -1, dtype=torch.int32, device=dev)
        if R_valid > 0:
            r2b, mask, rp = r2b[ray_mask_1], mask[ray_mask_1, :], rp[ray_mask_1].contiguous()
            loc_mask = torch.zeros([R_valid, S], dtype=torch.int32, device=dev)
            cum = torch.cumsum(mask, dim=-1).to(torch.int32)
            mask = (mask * cum * (cum <= S) - 1).contiguous()
            self.ext.get_shadingloc(rp, mask, R_valid, D, S, loc, loc_mask)
            radius = radius_limit_scale * max(self.voxel_size[0], self.voxel_size[1])
            self.ext.query_along_ray(self.points, r2b, R_valid, S, self.max_o, self.P, k, self.G, radius ** 2, self.shift,
                                     self.vdim, self.svs, self.ks, self.occ_numpnts, self.occ_2_pnts, self.coor_2_occ, loc,
                                     loc_mask, pidx)
            ray_mask_2 = (pidx.view(R_valid, -1) >= 0).sum(-1) > 0
            R_valid = int(ray_mask_2.sum())                                          # host sync (knnquery.py:277)
            loc, pidx = loc[ray_mask_2], pidx[ray_mask_2]
            ray_mask_1[ray_mask_1.clone()] = ray_mask_2
        return pidx, loc, ray_mask_1.view(B, R).to(torch.int8)


class ApiGrid:
    """Any torch_knnquery.VoxelGrid-compatible object behind the oracle's grid contract (``query_dense``), following the
    reference's own glue: set_pointset before every query (pointneus_disent.py:627, 353, 427), then
    utils.query (utils.py:90-113).  `sort_neighbours`: the reference kernel returns each slot's neighbour SET in
    arrival order; sorting it (by index) makes float sums over the K neighbours comparable between grids."""

    def __init__(self, voxel_grid, points: torch.Tensor, reinsert: bool = True, sort_neighbours: bool = True):
        self.vg, self.points, self.reinsert, self.sort = voxel_grid, points.reshape(1, -1, 3).contiguous(), reinsert, sort_neighbours
        self.n = torch.full((1,), self.points.shape[1], dtype=torch.int32, device=points.device)
        self.vg.set_pointset(self.points, self.n)

    def query_dense(self, raypos: torch.Tensor, k: int, radius_limit_scale: float, smax: int):
        if self.reinsert:
            self.vg.set_pointset(self.points, self.n)
        R = raypos.shape[0]
        pidx_v, loc_v, ray_mask = self.vg.query(raypos[None].contiguous(), k, radius_limit_scale, smax)
        rm = ray_mask.view(-1).bool()
        pidx = torch.full((R, smax, k), -1, dtype=torch.int32, device=raypos.device)
        loc = torch.zeros(R, smax, 3, dtype=raypos.dtype, device=raypos.device)
        if pidx_v.shape[0] > 0:
            if self.sort:   # -1 padding last, then ascending point id
                key = torch.where(pidx_v >= 0, pidx_v, torch.full_like(pidx_v, 2 ** 31 - 1))
                pidx_v = torch.gather(pidx_v, -1, key.argsort(-1))
            pidx[rm], loc[rm] = pidx_v, loc_v
        return {"pidx": pidx, "sample_loc": loc, "ray_mask2": rm.to(torch.int8)}


@contextmanager
def on_cuda():
    """The oracle builds its temporaries with plain factory calls (torch.zeros(...), torch.arange(...)): under this
    context they land on the GPU, so the restated graph runs there unchanged."""
    with torch.device("cuda"):
        yield


def params_to(P: H.Params, dev) -> H.Params:
    mv = lambda t: t.detach().to(dev).clone()
    P.neural_pts, P.neural_feats_color, P.neural_feats_geometry = mv(P.neural_pts), mv(P.neural_feats_color), mv(P.neural_feats_geometry)
    P.F_color = [(mv(W), mv(b)) for W, b in P.F_color]
    P.F_geometry = [(mv(W), mv(b)) for W, b in P.F_geometry]
    P.R = [(mv(W), mv(b)) for W, b in P.R]
    P.T = (mv(P.T[0]), mv(P.T[1]))
    P.beta = mv(P.beta)
    return P


def sample_z(P: H.Params, grid, uv, cam, training: bool, fast: int, rng=None):
    """The restated sampler (ray_sampler.py:377-588) alone, on the GPU."""
    with on_cuda():
        d, o = H.camera_rays(uv, cam["pose"], cam["intrinsics"])
        d = d.reshape(-1, 3)
        return H.sample_z(P, grid, d, o.unsqueeze(1).repeat(1, d.shape[0], 1).reshape(-1, 3), H.SamplerCfg(), training, fast, rng)


def make_grid(kind: str, P: H.Params, max_points_per_voxel: int = 128, max_occ_voxels: int = 32768, reinsert: bool = True):
    """kind = "reference": the compiled reference kernels; "product": spurfies_b200.knnquery.VoxelGrid -- same ctor
    arguments (pointneus_disent.py:45-62, caps above the occupancy so the reference's answer is well defined)."""
    a = P.grid_args
    if kind == "reference":
        vg = RefExtVoxelGrid(a["voxel_size"], a["voxel_scale"], a["kernel_size"], max_points_per_voxel, max_occ_voxels, a["ranges"])
    else:
        from spurfies_b200.knnquery import VoxelGrid
        vg = VoxelGrid(a["voxel_size"], a["voxel_scale"], a["kernel_size"], max_points_per_voxel, max_occ_voxels, a["ranges"])
    return ApiGrid(vg, P.neural_pts, reinsert=reinsert)


def training_step(P: H.Params, grid, uv, cam, rng, gt, with_tv: bool = True, z_vals=None):
    """Forward + VolSDFLoss + backward of the restated reference graph on the GPU (train.py:330-361 up to the optimiser).
    Returns (render outputs, loss dict); gradients are left in the .grad of P.trainable().  `z_vals` injects the sample
    depths (skipping the sampler) so that everything downstream can be compared at identical sample positions."""
    with on_cuda():
        out = H.render_forward(P, grid, uv, cam["pose"], cam["intrinsics"], H.SamplerCfg(), True, 1, rng, with_tv=with_tv,
                               z_vals=z_vals)
        lo = H.volsdf_loss(out, gt["rgb"], gt["mask"][0, :, 0])
        lo["loss"].backward()
    return out, lo
