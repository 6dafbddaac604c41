"""Drop-in mirrors of the reference model/renderer modules on the sm_100a kernels.

* ``PointVolSDF``            <- spurfies/model/pointneus_disent.py:24-908
* ``ErrorBoundSampler_pn``   <- spurfies/model/ray_sampler.py:337-588 (+ UniformSampler :17-59)
* ``LaplaceDensity``         <- spurfies/model/density.py:16-30
* ``VolSDFLoss``             <- spurfies/model/loss.py:19-100

Same method names, argument meaning, parameter names / shapes (reference checkpoints load unchanged:
``neural_pts``, ``neural_feats_color``, ``neural_feats_geometry``, ``F_color.{0,2,4,6}``, ``F_geometry.{0,2,4,6,8}``,
``T.0``, ``R.{0,2,4}``, ``density.beta``) and output dict keys.  Everything per-ray / per-sample / per-pair runs in
the hand-written kernels; torch is the plumbing (memory, streams, autograd bookkeeping, the scalar loss terms).

Differences from the reference, all deliberate: the voxel grid is built once and cached (the reference rebuilds it
3x per step); the random draws of the sampler can be injected (``rng=``) so that runs are reproducible against
the oracle; the constructor takes the neural points as tensors (reading the .ply is outside the hot path).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._lib import call, ptr, stream
from .fields import (ColorField, Composite, GeoPack, GeoSDF, LocalLoss, PseudoPointLoss, RadianceHead, SlotSet, TVRegul,
                     geo_sdf_raw,
                     local_feature_args, set_precision, surface_search)
from .knnquery import VoxelGrid


class _Conf(dict):
    """Duck-type of the pyhocon ConfigTree the reference passes around (get_int/get_float/... + attributes)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def _get(self, k, default=None):
        return self[k] if k in self else default

    get_int = get_float = get_bool = get_list = get_string = _get

    def get_config(self, k, default=None):
        v = self._get(k, default)
        return _Conf(v) if isinstance(v, dict) and not isinstance(v, _Conf) else v


def default_conf(near: float = 0.5, **over) -> _Conf:
    """config/vol/dtu_pn.yaml:21-44 merged with config/ours.yaml:21-23."""
    c = _Conf(feature_vector_size=64, scene_bounding_sphere=3.0, initialize_colors=True, k=8, r=2, rbf=45, vox_res=300,
              max_shading_pts=80, white_bkgd=False,
              density=_Conf(params_init=_Conf(beta=0.1), beta_min=0.0001),
              ray_sampler=_Conf(far=4.5, near=near, N_samples=64, N_samples_eval=128, N_samples_extra=32, eps=0.1,
                                beta_iters=10, max_total_iters=5))
    c.update(over)
    return c


_ZERO = {}


def _zero_scalar(dev) -> torch.Tensor:
    """A persistent read-only 0-d zero on `dev` (loss terms that are switched off): no fill launch per step."""
    dev = torch.device(dev)
    z = _ZERO.get(dev)
    if z is None:
        z = _ZERO[dev] = torch.zeros((), device=dev)
    return z


class LaplaceDensity(nn.Module):
    """alpha * Laplace(0, beta).cdf(-sdf)  (density.py:16-30).  The module form is API plumbing; on the hot path the
    density is evaluated inside the compositing / sampler kernels."""

    def __init__(self, params_init={}, beta_min=0.0001):
        super().__init__()
        for p in params_init:
            setattr(self, p, nn.Parameter(torch.tensor(float(params_init[p]))))
        self.register_buffer("beta_min", torch.tensor(float(beta_min)), persistent=False)

    def forward(self, sdf, beta=None):
        return self.density_func(sdf, beta=beta)

    def density_func(self, sdf, beta=None):
        if beta is None:
            beta = self.get_beta()
        alpha = 1 / beta
        return alpha * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))

    def get_beta(self):
        return self.beta.abs() + self.beta_min


class ErrorBoundSampler_pn:
    """VolSDF Algorithm 1 over the neural-point SDF (ray_sampler.py:337-588), one warp per ray in
    spf_sampler_iter.  ``get_z_vals`` keeps the reference signature; ``rng`` optionally injects the draws
    (t_rand [R,N_eval], u [R,N_samples], sampling_idx [N_extra]) that the reference takes from the global CPU
    generator (ray_sampler.py:55, 514, 550)."""

    def __init__(self, scene_bounding_sphere, near, far, N_samples, N_samples_eval, N_samples_extra, eps, beta_iters,
                 max_total_iters, inverse_sphere_bg=False, N_samples_inverse_sphere=0, add_tiny=0.0):
        if inverse_sphere_bg:
            raise NotImplementedError("inverse_sphere_bg is not used by any reference config")
        self.near = near
        self.far = 2.0 * scene_bounding_sphere  # ray_sampler.py:353 (the `far` argument is ignored there too)
        self.N_samples, self.N_samples_eval, self.N_samples_extra = N_samples, N_samples_eval, N_samples_extra
        self.eps, self.beta_iters, self.max_total_iters = eps, beta_iters, max_total_iters
        self.scene_bounding_sphere, self.add_tiny = scene_bounding_sphere, add_tiny
        self._const = {}

    def _consts(self, dev):
        if dev not in self._const:
            self._const[dev] = {
                "t_vals": torch.linspace(0.0, 1.0, steps=self.N_samples_eval).to(dev),
                "u_eval": torch.linspace(0.0, 1.0, steps=self.N_samples_eval).to(dev),
                "u_final": torch.linspace(0.0, 1.0, steps=self.N_samples).to(dev),
                # (1 / (4 log(1 + eps))) as the fp32 tensor arithmetic of ray_sampler.py:389 produces it
                "bound_coef": float(1.0 / (4.0 * torch.log(torch.tensor(self.eps + 1.0)))),
            }
        return self._const[dev]

    def get_z_vals(self, ray_dirs, cam_loc, model, fast=-1, iter_step=None, rng: Optional[Dict] = None):
        """ray_dirs [R,3], cam_loc [R,3] (all rows equal, as the reference builds it) or [3] -> (z [R,98], z_eik [R,1])."""
        dev = ray_dirs.device
        R = ray_dirs.shape[0]
        cst = self._consts(dev)
        max_total_iters = fast if fast >= 0 else self.max_total_iters
        ray_dirs = ray_dirs.contiguous().float()
        o = (cam_loc[0] if cam_loc.dim() == 2 else cam_loc).contiguous().float()
        training = model.training
        M = self.N_samples_eval
        t_rand = None
        if training:
            t_rand = (rng["t_rand"] if rng is not None else torch.rand(R, M)).to(dev, non_blocking=True).contiguous()
        z = torch.empty(R, M, dtype=torch.float32, device=dev)
        pts = torch.empty(R, M, 3, dtype=torch.float32, device=dev)
        call("spf_sampler_coarse", ptr(cst["t_vals"]), ptr(t_rand), float(self.near), float(self.far), ptr(o),
             ptr(ray_dirs), R, M, ptr(z), ptr(pts), stream())
        beta0 = model.density.get_beta().detach().reshape(1).float().contiguous()
        beta_io = torch.empty(R, dtype=torch.float32, device=dev)
        # device-side loop state of Algorithm 1 (ray_sampler.py:466-474): [0] = "some ray's beta still above beta0" in the
        # current iteration, [1] = "converged in an earlier iteration".  The reference decides with a host sync per
        # iteration (`beta.max() > beta0`, :468); here every launch of the multi-iteration (eval) schedule is predicated
        # on this state instead (spf_sampler_iter_pred), so a full-image render has no host round trip.
        state = torch.zeros(2, dtype=torch.int32, device=dev)
        # how many iterations of Algorithm 1 this batch used (device tensor, no sync; tests read it): the loop is
        # batch-global (ray_sampler.py:466-468), so a ray's samples depend on it
        iters_used = torch.ones(1, dtype=torch.int32, device=dev)
        self.last_iters_used = iters_used
        n_extra = self.N_samples_extra
        cols = self.N_samples + 2 + n_extra
        total_iters = 0
        sdf = None
        new_pts, new_z = pts, z
        z_out = p_out = None

        def final_draw(M, pred):
            nonlocal z_out, p_out
            u = None
            if training:
                u = (rng["u"] if rng is not None else torch.rand(R, self.N_samples)).to(dev, non_blocking=True).contiguous()
                sidx = rng["sampling_idx"] if rng is not None else torch.randperm(M)[:n_extra]
            else:
                # ray_sampler.py:552-553; cached on the device per M: no H2D copy per draw (CUDA-graph capturable)
                key = ("sidx", M, n_extra)
                if key not in cst:
                    cst[key] = torch.linspace(0, M - 1, n_extra).long().to(dev, dtype=torch.int32).contiguous()
                sidx = cst[key]
            sidx = sidx.to(dev, dtype=torch.int32).contiguous()
            if z_out is None:
                z_out = torch.empty(R, cols, dtype=torch.float32, device=dev)
                p_out = torch.empty(R, cols, 3, dtype=torch.float32, device=dev)
            args = (ptr(z), ptr(sdf.contiguous()), R, M, ptr(beta0), float(self.eps), int(self.beta_iters), cst["bound_coef"],
                    float(self.add_tiny), int(total_iters == 1), ptr(beta_io), 1, self.N_samples, ptr(u), ptr(cst["u_final"]),
                    float(self.near), float(self.far), ptr(sidx), n_extra, ptr(o), ptr(ray_dirs), ptr(z_out), ptr(p_out))
            if pred == 0:
                call("spf_sampler_iter", *args, ptr(state), stream())
            else:
                call("spf_sampler_iter_pred", *args, ptr(state), pred, stream())

        while total_iters < max_total_iters:
            with torch.no_grad():
                # once every ray of the batch has converged (state[1], set on the device) the remaining iterations of
                # the predicated schedule read nothing of this: skip the search and the MLP on the device as well
                done = state[1:2] if (total_iters >= 1 and hasattr(model, "_point_slots")) else None
                s_new = (model.sdf_importance(new_pts.view(-1, 3), skip=done) if done is not None
                         else model.sdf_importance(new_pts.view(-1, 3))).view(R, -1)
            if sdf is None:
                sdf = s_new
            else:  # merge by the sort permutation (ray_sampler.py:405-415, 533)
                Mn = z.shape[1] + new_z.shape[1]
                z2 = torch.empty(R, Mn, dtype=torch.float32, device=dev)
                sdf2 = torch.empty(R, Mn, dtype=torch.float32, device=dev)
                call("spf_sampler_merge", ptr(z), ptr(sdf), z.shape[1], ptr(new_z), ptr(s_new.contiguous()),
                     new_z.shape[1], R, ptr(z2), ptr(sdf2), stream())
                z, sdf = z2, sdf2
            M = z.shape[1]
            total_iters += 1
            if total_iters >= max_total_iters:
                # the last allowed iteration samples either way (ray_sampler.py:466); single-iteration (training)
                # schedule: a plain launch
                final_draw(M, 0 if max_total_iters == 1 else 3)
                break
            # probe: line search for beta, error-bound opacity, N_samples_eval new samples per ray; sets state[0] if
            # any ray is still above beta0
            state[0:1].zero_()
            probe_z = torch.empty(R, self.N_samples_eval, dtype=torch.float32, device=dev)
            probe_p = torch.empty(R, self.N_samples_eval, 3, dtype=torch.float32, device=dev)
            beta_probe = beta_io.clone()
            call("spf_sampler_iter_pred", ptr(z), ptr(sdf.contiguous()), R, M, ptr(beta0), float(self.eps),
                 int(self.beta_iters), cst["bound_coef"], float(self.add_tiny), int(total_iters == 1),
                 ptr(beta_probe), 0, self.N_samples_eval, None, ptr(cst["u_eval"]), float(self.near),
                 float(self.far), None, 0, ptr(o), ptr(ray_dirs), ptr(probe_z), ptr(probe_p), ptr(state), 1, stream())
            final_draw(M, 2)                                                  # runs only if this iteration converged
            state[1:2].bitwise_or_((state[0:1] == 0).to(torch.int32))         # converged now or earlier
            iters_used.add_((state[1:2] == 0).to(torch.int32))                # another iteration's result will be used
            beta_io, new_z, new_pts = beta_probe, probe_z, probe_p
        if z_out is not None:
            self.last_points = p_out
        if z_out is None:  # max_total_iters == 0: the reference would fail on `samples`; return the coarse samples
            z_out, self.last_points = z, pts
        if rng is not None and "eik_idx" not in rng:
            z_samples_eik = z_out[:, :1]  # ray_sampler.py:561-563 draws a random column; nothing downstream reads it
        else:
            idx = rng["eik_idx"].to(dev) if rng is not None else torch.randint(z_out.shape[-1], (R,), device=dev)
            z_samples_eik = torch.gather(z_out, 1, idx.unsqueeze(-1))
        return z_out, z_samples_eik


class PointVolSDF(nn.Module):
    _instances = 0

    def __init__(self, conf, scan_id=None, dataset=None, neural_points: Optional[torch.Tensor] = None,
                 neural_colors: Optional[torch.Tensor] = None, device="cuda", ranges=None,
                 max_points_per_voxel: int = 26, max_occ_voxels: int = 20000, precision: str = "fp32"):
        super().__init__()
        # "fp32": exact SIMT kernels (1e-4 vs the reference); "bf16": tcgen05 tensor-core kernels (2e-2)
        assert precision in ("fp32", "bf16")
        self.precision = precision
        if not isinstance(conf, _Conf) and isinstance(conf, dict):
            conf = _Conf(conf)
        self.conf = conf
        self.scan_id, self.dataset = scan_id, dataset
        self.feature_vector_size = conf.get_int("feature_vector_size")
        self.scene_bounding_sphere = conf.get_float("scene_bounding_sphere", default=1.0)
        self.white_bkgd = conf.get_bool("white_bkgd", default=False)
        self.register_buffer("bg_color", torch.tensor(conf.get_list("bg_color", default=[1.0, 1.0, 1.0])).float(),
                             persistent=False)
        self.conf.rbf = 45  # pointneus_disent.py:42
        if ranges is None:  # pointneus_disent.py:45-62
            big = str(scan_id) in ("garden", "stump") and dataset == "mipnerf"
            ranges = (-2, -2, -2, 2, 2, 2) if big else (-1, -1, -1, 1, 1, 1)
        self._voxel_grid_neural = VoxelGrid((0.025, 0.025, 0.025), (3, 3, 3), (3, 3, 3), max_points_per_voxel,
                                            max_occ_voxels, ranges)
        if neural_points is None:
            # pointneus_disent.py:131-146: the scene's DUSt3R cloud, voxel-downsampled at conf.vox_res
            import os
            from .ingest import load_neural_points
            path = conf.get_string("pointcloud_path", default=None)
            if path is None:
                if dataset == "dtu":
                    path = f"./data/{dataset}/scan{scan_id}/{scan_id}.ply"
                elif dataset in ("mipnerf", "own_data"):
                    path = f"./data/{dataset}/{scan_id}/{scan_id}.ply"
                else:
                    raise NotImplementedError(dataset)
            if not os.path.exists(path):
                raise RuntimeError(f"The pointcloud_path must be specified ({path} not found); or pass `neural_points`")
            data = load_neural_points(path, vox_res=conf.get_int("vox_res", default=None))
            neural_points, neural_colors = data["pts"].float(), (data["colors"].float() if "colors" in data else None)
        self._init_neural_info(neural_points, neural_colors, device)
        C = conf.feature_vector_size
        self.F_color = nn.Sequential(nn.Linear(C + 39, 256), nn.LeakyReLU(inplace=True), nn.Linear(256, 256),
                                     nn.LeakyReLU(inplace=True), nn.Linear(256, 256), nn.LeakyReLU(inplace=True),
                                     nn.Linear(256, 256))
        self.F_geometry = nn.Sequential(nn.Linear(C // 2 + 3, 256), nn.LeakyReLU(inplace=True), nn.Linear(256, 256),
                                        nn.LeakyReLU(inplace=True), nn.Linear(256, 256), nn.LeakyReLU(inplace=True),
                                        nn.Linear(256, 256), nn.LeakyReLU(inplace=True), nn.Linear(256, 256))
        self.T = nn.Sequential(nn.Linear(256, 1))
        self.R = nn.Sequential(nn.Linear(256 + 21, 256), nn.LeakyReLU(inplace=True), nn.Linear(256, 256),
                               nn.LeakyReLU(inplace=True), nn.Linear(256, 3), nn.Sigmoid())
        self.density = LaplaceDensity(**conf.get_config("density"))
        self.ray_sampler = ErrorBoundSampler_pn(self.scene_bounding_sphere, **conf.get_config("ray_sampler"))
        self._geo_pack = GeoPack()
        self._self_knn = None
        self._dp = (1, None)   # (world size, process group) of the ray-sharded data parallelism, see set_data_parallel
        PointVolSDF._instances += 1
        self._owner = f"m{PointVolSDF._instances}."   # prefix of this model's arena buffers (fields.Arena)
        self.to(device)

    # ------------------------------------------------------------------ parameters (pointneus_disent.py:110-205)
    def _init_neural_info(self, pts, colors, device):
        n = len(pts)
        C = self.conf.feature_vector_size
        self.register_buffer("neural_pts", pts.detach().float().clone().contiguous())
        fc = torch.empty(n, C).uniform_(-1e-4, 1e-4)
        fg = torch.empty(n, C // 2).normal_(0.0, 0.01)
        norms = fg.norm(dim=-1, keepdim=True)
        fg = fg * (torch.clamp(norms, max=1) / (norms + 1e-7))
        if colors is not None and self.conf.get_bool("initialize_colors", default=True):
            fc[:, :3] = colors.float().to(fc.device) * 2.0 / 255.0 - 1.0
        self.neural_feats_color = nn.Parameter(fc)
        self.neural_feats_geometry = nn.Parameter(fg)

    def set_data_parallel(self, world_size: int, group=None) -> None:
        """Ray-sharded data parallelism (new functionality): with world_size > 1 the training forward
        all-reduces the denominators of its count-normalised loss terms (pseudo-point, local, eikonal) so that the
        gradient averaged over ranks is the gradient of the one-big-batch step.  The ray shards must be equally
        sized for the per-ray means (rgb, mask) to combine exactly."""
        self._dp = (int(world_size), group)

    # ------------------------------------------------------------------ helpers
    def _grid(self) -> VoxelGrid:
        # the reference re-inserts the points before every query (pointneus_disent.py:627, 353, 427, 252);
        # set_pointset caches on (data_ptr, version) so this is free after the first call
        if getattr(self, "_n_tensor", None) is None or self._n_tensor.device != self.neural_pts.device:
            self._n_tensor = torch.full((1,), len(self.neural_pts), device=self.neural_pts.device, dtype=torch.int)
        self._voxel_grid_neural.set_pointset(self.neural_pts.unsqueeze(0), self._n_tensor)
        return self._voxel_grid_neural

    def _pack(self) -> GeoPack:
        return self._geo_pack.get(self.F_geometry, self.T)

    def _point_slots(self, x: torch.Tensor, tag: str = "points", skip: Optional[torch.Tensor] = None) -> SlotSet:
        pidx = self._grid().query_points(x.contiguous().float(), self.conf.k, self.conf.r, skip=skip)
        return SlotSet(pidx, tag, self._owner)

    # ------------------------------------------------------------------ point SDF queries
    def sdf_importance(self, inputs: torch.Tensor, skip: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pointneus_disent.py:348-421: SDF at points [N,3] -> [N], 1000 where no neighbour.  ``skip`` (device int32
        [1], optional): non-zero at launch time = the caller will not read the result (VoxelGrid.query_points): every
        value is the 1000 filler and neither the search nor the MLP runs."""
        set_precision(self.precision)
        x = inputs.detach().contiguous().float()
        slots = self._point_slots(x, skip=skip)
        sdf, _, _ = geo_sdf_raw(self._pack(), slots, x, self.neural_pts, self.neural_feats_geometry.detach(),
                                self.conf.rbf, False, False)
        return sdf

    def get_sdf_eval(self, inputs: torch.Tensor) -> torch.Tensor:
        """pointneus_disent.py:249-298 (same arithmetic as sdf_importance)."""
        return self.sdf_importance(inputs)

    def pseudo_sdf(self, inputs: torch.Tensor, dense: bool = False):
        """pointneus_disent.py:423-495: SDF at points with autograd; [V,1] over the valid points (reference
        contract, one host sync) or, with dense=True, ([N] with 1000 fill, valid mask) without a sync."""
        set_precision(self.precision)
        x = inputs.contiguous().float()
        slots = self._point_slots(x.detach(), "pseudo")
        sdf, _ = GeoSDF.apply(self.neural_feats_geometry, x, slots, self._pack(), self.neural_pts, self.conf.rbf, False)
        if dense:
            return sdf, slots.valid_mask()
        if slots.V == 0:
            return torch.ones(x.shape[0], device=x.device) * 1000
        return sdf[slots.list[:slots.V].long()].unsqueeze(-1)

    def find_surface_points(self, sdf, d_all, device="cuda"):
        """pointneus_disent.py:586-612: sdf, d_all [..., Rv, S] (1000 = slot without neighbours) -> (d_surface, mask)
        [..., Rv]: depth of the first back-facing zero crossing per ray.  (The reference also overwrites the 1000s of
        its argument with NaN in place; nothing downstream reads that.)"""
        shape = sdf.shape[:-1]
        S = sdf.shape[-1]
        sd = sdf.detach().reshape(-1, S).float().contiguous()
        dd = d_all.detach().reshape(-1, S).float().contiguous()
        Rv = sd.shape[0]
        zero3 = torch.zeros(3, device=sd.device)
        _, cross, d_surface, _, _ = surface_search(sd, dd, zero3, torch.zeros(max(Rv, 1), 3, device=sd.device), Rv, S)
        return d_surface.reshape(shape), (cross >= 0).reshape(shape)

    def volume_rendering(self, deltas, density):
        """pointneus_disent.py:894-908 (API compatibility; the hot path uses the fused compositing kernel)."""
        free_energy = deltas * density
        shifted = torch.cat([torch.zeros(deltas.shape[0], 1, device=deltas.device), free_energy[:, :-1]], dim=-1)
        return (1 - torch.exp(-free_energy)) * torch.exp(-torch.cumsum(shifted, dim=-1))

    def tv_loss(self):
        """tv_regul(neural_pts, neural_feats_geometry) (utils.py:221-281); the self-kNN lists are cached because
        neural_pts is a buffer that never changes (the reference recomputes them every step)."""
        key = (self.neural_pts.data_ptr(), self.neural_pts._version)
        if self._self_knn is None or self._self_knn[0] != key:
            self._self_knn = (key, self._grid().query_points(self.neural_pts, self.conf.k, self.conf.r))
        world, group = self._dp
        if world > 1 and self.training:
            # ray-independent term: every rank evaluates 1/world of the points, scaled by world, so that the gradient
            # average over the ranks is the whole regulariser (each rank reports its share x world as the term's value)
            import torch.distributed as dist
            from .dist import shard_range
            lo, hi = shard_range(self.neural_pts.shape[0], dist.get_rank(group), world)
            return TVRegul.apply(self.neural_feats_geometry, self.neural_pts, self._self_knn[1], lo, hi - lo, float(world))
        return TVRegul.apply(self.neural_feats_geometry, self.neural_pts, self._self_knn[1])

    # ------------------------------------------------------------------ forward (pointneus_disent.py:614-892)
    def forward(self, input, fast=-1, rng=None, dense_outputs: bool = False, aux_losses: bool = True):
        """aux_losses=False skips the pseudo-point and TV terms (pointneus_disent.py:765-780, 870-878), which the
        reference also computes in eval mode although nothing reads them there (eval_spurfies.py:278-290)."""
        set_precision(self.precision)
        intrinsics, uv, pose = input["intrinsics"], input["uv"], input["pose"]
        iter_step = input.get("iter_step", 1)
        dev = self.neural_pts.device
        R = uv.shape[1]
        K, S = self.conf.k, self.conf.max_shading_pts
        grid = self._grid()
        # rays (rend_util.py:60-95)
        ray_dirs = torch.empty(R, 3, dtype=torch.float32, device=dev)
        cam_loc = torch.empty(3, dtype=torch.float32, device=dev)
        depth_scale = torch.empty(R, dtype=torch.float32, device=dev)
        call("spf_camera_rays", ptr(uv.reshape(-1, 2).float().contiguous()), ptr(pose.reshape(4, 4).float().contiguous()),
             ptr(intrinsics.reshape(4, 4).float().contiguous()), R, ptr(ray_dirs), ptr(cam_loc), ptr(depth_scale),
             stream())
        # importance sampling along the rays (pointneus_disent.py:647-649)
        z_vals, _ = self.ray_sampler.get_z_vals(ray_dirs, cam_loc, self, fast, iter_step, rng=rng)
        points = self.ray_sampler.last_points  # cam_loc + z * dir, produced by the sampler kernel
        # kNN (pointneus_disent.py:654-660)
        pidx, loc, slot_sample, nvalid = grid.query_dense(points, K, self.conf.r, S)
        slots = SlotSet(pidx, "fine", self._owner)
        n = R * S
        # filter_points (pointneus_disent.py:666-669)
        t = torch.empty(R, S, dtype=torch.float32, device=dev)
        delta = torch.empty(R, S, dtype=torch.float32, device=dev)
        x_new = torch.empty(R, S, 3, dtype=torch.float32, device=dev)
        call("spf_ray_prep", ptr(loc), ptr(pidx), ptr(cam_loc), ptr(ray_dirs), R, S, K, ptr(t), ptr(delta), ptr(x_new),
             stream())
        xs = x_new.view(n, 3)
        # fields
        sdf, grad = GeoSDF.apply(self.neural_feats_geometry, xs, slots, self._pack(), self.neural_pts, self.conf.rbf,
                                 True)
        fc = [m for m in self.F_color if isinstance(m, nn.Linear)]
        rl = [m for m in self.R if isinstance(m, nn.Linear)]
        hbar = ColorField.apply(self.neural_feats_color, fc[0].weight, fc[0].bias, fc[1].weight, fc[1].bias,
                                fc[2].weight, fc[2].bias, xs, slots, self.neural_pts, self.conf.rbf)
        slots.single_consumer = True   # hbar feeds only the radiance head: its gradient may travel in compact bf16 form
        rgb_s = RadianceHead.apply(hbar, fc[3].weight, fc[3].bias, rl[0].weight, rl[0].bias, rl[1].weight, rl[1].bias,
                                   rl[2].weight, rl[2].bias, ray_dirs, slots, S)
        beta = self.density.get_beta()
        weights, rgb, depth, acc, dist, normal = Composite.apply(sdf, rgb_s, beta, delta.view(-1), t.view(-1), grad,
                                                                 slots.pidx, nvalid, R, S, K, not self.training)
        ray_mask = nvalid > 0
        # feature-consistency loss at the first back-facing zero crossing (pointneus_disent.py:727-763, DTU 3-view only)
        local_data = input.get("local_data", None)
        local_loss = _zero_scalar(dev)
        world, group = self._dp
        dp = world > 1 and self.training
        counts = {}
        if local_data is not None and self.training:
            feats = local_feature_args(local_data, dev)
            local_loss, d_surface, cross = LocalLoss.apply(sdf, t, cam_loc, ray_dirs, feats, R, S)
            self._last_surface = (d_surface, cross)
            if dp:
                counts["local"] = (cross >= 0).sum()
        # pseudo points (pointneus_disent.py:765-780): expected-depth point per hit ray, SDF should vanish there
        if aux_losses:
            aux = {} if dp else None
            pseudo = PseudoPointLoss.apply(self.neural_feats_geometry, dist, cam_loc, ray_dirs, nvalid, grid, self.conf.k,
                                           self.conf.r, self._pack(), self.neural_pts, self.conf.rbf, self._owner, aux)
            if dp:
                counts["pseudo"] = aux["pseudo_count"]
        else:
            pseudo = _zero_scalar(dev)
        eik_scale = None
        if dp:
            # global-count normalisation of the count-normalised means (one 12-byte all-reduce, no host sync)
            from .dist import global_count_scales
            zero_c = torch.zeros((), dtype=torch.int64, device=dev)
            local_counts = torch.stack([counts.get("pseudo", zero_c).to(torch.int64),
                                        counts.get("local", zero_c).to(torch.int64), slots.count.reshape(()).to(torch.int64)])
            sc = global_count_scales(local_counts, world, group)
            pseudo, local_loss, eik_scale = pseudo * sc[0], local_loss * sc[1], sc[2]
        far_cfg = float(self.conf.ray_sampler.far)
        depth_vals = torch.where(ray_mask[:, None], t * depth_scale[:, None], torch.full_like(t, far_cfg))
        output = {
            "rgb_values": rgb,
            "depth_values": depth.unsqueeze(-1),
            "depth_vals": depth_vals,
            "weights": weights,
            "xyz": x_new,
            "local_loss": local_loss,
            "pseudo_pts_loss": pseudo,
            "tv_loss": self.tv_loss() if aux_losses else _zero_scalar(dev),
        }
        if self.white_bkgd:  # pointneus_disent.py:856-861 (computed but never written back there either)
            pass
        if eik_scale is not None:
            output["eikonal_scale"] = eik_scale
        if not self.training:
            output["normal_map"] = normal
        else:
            if dense_outputs:
                output["grad_theta_dense"], output["grad_theta_mask"] = grad, slots.valid_mask()
            else:
                output["grad_theta"] = grad[slots.list[:slots.V].long()]
        # detached views only: holding graph tensors here would keep last step's autograd graph (and its
        # AccumulateGrad nodes) alive, which breaks CUDA-graph capture of the next step
        self._last = {"slots": slots, "z_vals": z_vals, "ray_mask": ray_mask, "sdf": sdf.detach(), "delta": delta, "t": t,
                      "rgb_s": rgb_s.detach(), "dist": dist.detach(), "acc": acc.detach(), "ray_dirs": ray_dirs,
                      "cam_loc": cam_loc, "loc": loc, "slot_sample": slot_sample}
        return output


class VolSDFLoss(nn.Module):
    """loss.py:19-100.  Accepts the reference's ragged ``grad_theta`` or the dense pair
    (``grad_theta_dense``, ``grad_theta_mask``) so that a training step needs no host sync."""

    def __init__(self, rgb_loss="torch.nn.L1Loss", local_weight=0.5, pseudo_weight=0.5, eikonal_weight=0.001,
                 rgb_weight=1.0, tv_weight=0.01):
        super().__init__()
        self.local_weight, self.pseudo_weight, self.eikonal_weight = local_weight, pseudo_weight, eikonal_weight
        self.rgb_weight, self.tv_weight = rgb_weight, tv_weight
        self.rgb_loss = nn.L1Loss(reduction="mean")
        self.iter_step = 0

    def forward(self, model_outputs, ground_truth):
        dev = model_outputs["rgb_values"].device
        if dev.type == "cuda" and "grad_theta" not in model_outputs and "weights" in model_outputs:
            return self._fused(model_outputs, ground_truth)
        rgb_gt = ground_truth["rgb"].to(dev).reshape(-1, 3)
        mask_gt = ground_truth["mask"].to(dev)
        zero = torch.zeros((), device=dev)
        out = {"rgb_loss": self.rgb_loss(model_outputs["rgb_values"], rgb_gt)}
        if "grad_theta" in model_outputs:
            out["eikonal_loss"] = ((model_outputs["grad_theta"].norm(2, dim=1) - 1) ** 2).mean()
        elif "grad_theta_dense" in model_outputs:
            g, m = model_outputs["grad_theta_dense"], model_outputs["grad_theta_mask"]
            out["eikonal_loss"] = (((g.norm(2, dim=1) - 1) ** 2) * m).sum() / m.sum().clamp(min=1)
        else:
            out["eikonal_loss"] = zero
        if "eikonal_scale" in model_outputs:   # data parallel: this rank's share of the global mean (PointVolSDF.forward)
            out["eikonal_loss"] = out["eikonal_loss"] * model_outputs["eikonal_scale"]
        out["tv_loss"] = model_outputs["tv_loss"] if ("tv_loss" in model_outputs and self.tv_weight > 0) else zero
        if "weights" in model_outputs:
            wsum = model_outputs["weights"].sum(-1, keepdim=True)
            out["mask_loss"] = F.binary_cross_entropy(wsum.clip(1e-3, 1.0 - 1e-3), mask_gt.squeeze()[:, 0][..., None])
        else:
            out["mask_loss"] = zero
        out["local_loss"] = model_outputs.get("local_loss", zero)
        out["pseudo_loss"] = model_outputs["pseudo_pts_loss"] if ("pseudo_pts_loss" in model_outputs and self.pseudo_weight > 0) else zero
        out["loss"] = (self.rgb_weight * out["rgb_loss"] + self.eikonal_weight * out["eikonal_loss"]
                       + self.tv_weight * out["tv_loss"] + self.local_weight * out["local_loss"]
                       + self.pseudo_weight * out["pseudo_loss"] + out["mask_loss"])
        self.iter_step += 1
        return out

    def _fused(self, model_outputs, ground_truth):
        """The same terms from spf_volsdf_loss (two launches instead of ~70 torch kernels); used for the dense,
        sync-free outputs of ``PointVolSDF.forward(dense_outputs=True)`` / eval outputs.  The ragged ``grad_theta`` of the
        reference contract takes the torch path above."""
        dev = model_outputs["rgb_values"].device
        rgb_gt = ground_truth["rgb"].to(dev).reshape(-1, 3).float().contiguous()
        mask_gt = ground_truth["mask"].to(dev).float()
        mask2 = mask_gt.squeeze()
        mask2 = mask2.reshape(mask2.shape[0], -1).contiguous()          # loss.py:83: mask.squeeze()[:, 0]
        zero = _zero_scalar(dev)
        tv = model_outputs.get("tv_loss") if self.tv_weight > 0 else None
        pseudo = model_outputs.get("pseudo_pts_loss") if self.pseudo_weight > 0 else None
        local = model_outputs.get("local_loss")
        g = model_outputs.get("grad_theta_dense")
        m = model_outputs.get("grad_theta_mask") if g is not None else None
        wts = (self.rgb_weight, self.eikonal_weight, self.tv_weight, self.local_weight, self.pseudo_weight)
        loss, terms = _FusedLoss.apply(model_outputs["rgb_values"], model_outputs["weights"],
                                       tv if tv is not None else zero, local if local is not None else zero,
                                       pseudo if pseudo is not None else zero, rgb_gt, mask2, g, m, wts)
        es = model_outputs.get("eikonal_scale")
        if es is not None and g is not None:   # data parallel: the eikonal mean over the GLOBAL valid-sample count (value only:
            eik = terms[2]                     # the term has no gradient path)
            loss = loss + self.eikonal_weight * (es - 1.0) * eik
            terms = terms.clone()
            terms[2] = eik * es
            terms[0] = loss.detach()
        self.iter_step += 1
        return {"loss": loss, "rgb_loss": terms[1], "eikonal_loss": terms[2], "tv_loss": terms[3], "mask_loss": terms[4],
                "local_loss": terms[5], "pseudo_loss": terms[6]}


class _FusedLoss(torch.autograd.Function):
    """loss.py:51-100 through spf_volsdf_loss.  Differentiable in rgb_values, weights and the three scalar terms; the
    eikonal term has no gradient path (grad_theta is first-order only)."""

    @staticmethod
    def forward(ctx, rgb, weights, tv, local, pseudo, rgb_gt, mask2, grad_dense, grad_mask, wts):
        dev = rgb.device
        R, S = weights.shape
        f32 = lambda v: v.detach().float().contiguous()
        rgb_c, w_c = f32(rgb), f32(weights)
        sc = [f32(v).reshape(1) for v in (tv, local, pseudo)]
        terms = torch.empty(8, dtype=torch.float32, device=dev)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        d_rgb = torch.empty(R, 3, dtype=torch.float32, device=dev) if need else None
        d_w = torch.empty(R, S, dtype=torch.float32, device=dev) if need else None
        gd = f32(grad_dense) if grad_dense is not None else None
        gm = grad_mask.contiguous().view(torch.uint8) if grad_mask is not None else None
        n = gd.shape[0] if gd is not None else 0
        from .fields import Arena
        ws = Arena.get("loss_ws", (_lib.lib.spf_loss_workspace_bytes(),), torch.uint8, dev)
        call("spf_volsdf_loss", ptr(rgb_c), ptr(rgb_gt), ptr(w_c), ptr(mask2), int(mask2.shape[1]), ptr(gd), ptr(gm), n, R, S,
             ptr(sc[0]), ptr(sc[1]), ptr(sc[2]), *[float(v) for v in wts], ptr(terms), ptr(d_rgb), ptr(d_w), ptr(ws),
             ws.numel(), stream())
        ctx.saved_t = (d_rgb, d_w)
        ctx.wts = wts
        ctx.mark_non_differentiable(terms)
        return terms[0], terms

    @staticmethod
    def backward(ctx, g, _):
        d_rgb, d_w = ctx.saved_t
        _, _, w_tv, w_local, w_pseudo = ctx.wts
        return (d_rgb * g if d_rgb is not None else None, d_w * g if d_w is not None else None, g * w_tv, g * w_local,
                g * w_pseudo, None, None, None, None, None)
