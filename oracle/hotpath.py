"""oracle/hotpath.py -- TEST INFRASTRUCTURE ONLY (CPU checker; never imported by the product).

Pure-torch, device-agnostic, RNG-injectable restatement of the reference per-ray hot path
(kevinYitshak/spurfies @ 858a95f).  Every function cites the reference lines it follows.
It is validated against the reference's own Python modules imported in the authoring
container (tests/golden/make_golden.py -> tests/golden/*.pt, checked in
tests/test_oracle_hotpath.py), and is the CPU baseline timed by bench.py.

Conventions: all tensors fp32; kNN comes from oracle.knn.OracleGrid (C restatement of the
reference kernels); stock torch ops (nn.functional.linear, index_add_, cumsum, searchsorted,
sort) are used exactly where the reference uses them.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .knn import OracleGrid

LEAKY = 0.01  # nn.LeakyReLU default slope (pointneus_disent.py:76-107)
# "grid": C restatement of the reference voxel-grid kernels; "cdist": the reference test's pure-torch brute force
# (torch.cdist + topk), which is what bench.py times as the CPU baseline.
KNN_BACKEND = "grid"


# ----------------------------------------------------------------------------- parameters
@dataclass
class Params:
    """Same names / shapes as the reference state_dict (SURVEY section 5 checkpoint row)."""
    neural_pts: torch.Tensor                       # [N,3] buffer
    neural_feats_color: torch.Tensor               # [N,64]
    neural_feats_geometry: torch.Tensor            # [N,32]
    F_color: list                                  # [(W[256,103],b), (W[256,256],b) x3]
    F_geometry: list                               # [(W[256,35],b), (W[256,256],b) x4]
    T: tuple                                       # (W[1,256], b[1])
    R: list                                        # [(W[256,277],b),(W[256,256],b),(W[3,256],b)]
    beta: torch.Tensor                             # scalar parameter (density.beta)
    beta_min: float = 1e-4
    rbf: float = 45.0                              # pointneus_disent.py:42
    k: int = 8
    r: float = 2.0
    max_shading_pts: int = 80
    grid_args: dict = field(default_factory=lambda: dict(
        voxel_size=(0.025, 0.025, 0.025), voxel_scale=(3, 3, 3), kernel_size=(3, 3, 3),
        ranges=(-1, -1, -1, 1, 1, 1)))             # pointneus_disent.py:45-62

    def trainable(self):
        out = [self.neural_feats_color, self.neural_feats_geometry, self.beta]
        for W, b in self.F_color + self.R:
            out += [W, b]
        return out

    def make_grid(self) -> OracleGrid:
        return OracleGrid(self.neural_pts, **self.grid_args)


def init_params(neural_pts: torch.Tensor, colors: Optional[torch.Tensor] = None, seed: int = 0,
                feature_vector_size: int = 64, **kw) -> Params:
    """Random-init parameters the way the reference constructor does
    (pointneus_disent.py:76-107 default nn.Linear init; :117-129, 183-199 latent init)."""
    g = torch.Generator().manual_seed(seed)
    torch_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        def lin(i, o):
            m = torch.nn.Linear(i, o, bias=True)
            return (m.weight.detach().clone(), m.bias.detach().clone())
        C, G = feature_vector_size, feature_vector_size // 2
        F_color = [lin(C + 39, 256), lin(256, 256), lin(256, 256), lin(256, 256)]
        F_geometry = [lin(G + 3, 256), lin(256, 256), lin(256, 256), lin(256, 256), lin(256, 256)]
        T = lin(256, 1)
        R = [lin(256 + 21, 256), lin(256, 256), lin(256, 3)]
        N = neural_pts.shape[0]
        fc = torch.empty(N, C).uniform_(-1e-4, 1e-4)
        fg = torch.empty(N, G).normal_(0.0, 0.01)
        norms = fg.norm(dim=-1, keepdim=True)
        fg = fg * (torch.clamp(norms, max=1) / (norms + 1e-7))
        if colors is not None:
            fc[:, :3] = colors.float() * 2.0 / 255.0 - 1.0
    finally:
        torch.random.set_rng_state(torch_state)
    del g
    return Params(neural_pts=neural_pts.float().contiguous(), neural_feats_color=fc, neural_feats_geometry=fg,
                  F_color=F_color, F_geometry=F_geometry, T=T, R=R, beta=torch.tensor(0.1), **kw)


# ----------------------------------------------------------------------------- small pieces
def positional_encoding(x: torch.Tensor, multires: int) -> torch.Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]  (embedder.py:10-36)."""
    out = [x]
    for freq in 2.0 ** torch.linspace(0.0, multires - 1, multires):
        out += [torch.sin(x * freq), torch.cos(x * freq)]
    return torch.cat(out, -1)


def mlp(x, layers, act_last=False):
    n = len(layers)
    for i, (W, b) in enumerate(layers):
        x = F.linear(x, W, b)
        if i < n - 1 or act_last:
            x = F.leaky_relu(x, LEAKY)
    return x


# Arithmetic model of the product's tensor-core ("bf16") mode for the trainable colour path, a DIAGNOSTIC for that mode
# (tools/bf16_grad_study.py, printed by the bf16 test; parity itself is asserted against the reference golden): operands
# of every FORWARD matrix product (inputs, weights, inter-layer activations) are rounded to fp16 -- the kernels' forward
# operand format, csrc/umma.cuh -- products are accumulated in fp32, biases / LeakyReLU / interpolation / sigmoid stay fp32; the per-ray PE3(dir) columns of R.0 stay
# fp32; F_color.6 (linear, no activation: pointneus_disent.py:83) is applied after the neighbour interpolation.
# Off by default: the oracle then follows the reference's fp32 graph op for op.
BF16_OPERANDS = False


def _bf(t):
    return t.to(torch.float16).float()  # differentiable (identity gradient)


def _color_bf16_operands(p: Params, fin, dirs_v, w, norm, idx, V):
    (W1, b1), (W2, b2), (W3, b3), (W4, b4) = p.F_color
    (R1, rb1), (R2, rb2), (R3, rb3) = p.R
    h = F.leaky_relu(F.linear(_bf(fin), _bf(W1), b1), LEAKY)
    h = F.leaky_relu(F.linear(_bf(h), _bf(W2), b2), LEAKY)
    h = F.leaky_relu(F.linear(_bf(h), _bf(W3), b3), LEAKY)
    hbar = torch.zeros(V, 256).index_add_(0, idx, (w / norm[idx])[:, None] * h)
    f = F.linear(_bf(hbar), _bf(W4), b4)
    zpe = F.linear(positional_encoding(dirs_v, 3), R1[:, :21], rb1)
    a1 = F.leaky_relu(F.linear(_bf(f), _bf(R1[:, 21:])) + zpe, LEAKY)
    a2 = F.leaky_relu(F.linear(_bf(a1), _bf(R2), rb2), LEAKY)
    return torch.sigmoid(F.linear(_bf(a2), _bf(R3), rb3))


def get_beta(p: Params):
    return p.beta.abs() + p.beta_min  # density.py:28-30


def laplace_density(sdf, beta):
    """density.py:21-26"""
    alpha = 1 / beta
    return alpha * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))


def volume_rendering(deltas, density):
    """pointneus_disent.py:894-908"""
    free_energy = deltas * density
    shifted = torch.cat([torch.zeros(deltas.shape[0], 1), free_energy[:, :-1]], dim=-1)
    alpha = 1 - torch.exp(-free_energy)
    transmittance = torch.exp(-torch.cumsum(shifted, dim=-1))
    return alpha * transmittance


def camera_rays(uv, pose, intrinsics):
    """rend_util.py:60-95, 143-156 (pose-matrix branch). uv [1,R,2] -> dirs [1,R,3], cam_loc [1,3]."""
    cam_loc = pose[:, :3, 3]
    fx, fy = intrinsics[:, 0, 0], intrinsics[:, 1, 1]
    cx, cy, sk = intrinsics[:, 0, 2], intrinsics[:, 1, 2], intrinsics[:, 0, 1]
    x, y = uv[:, :, 0], uv[:, :, 1]
    z = torch.ones_like(x)
    x_lift = (x - cx[:, None] + cy[:, None] * sk[:, None] / fy[:, None] - sk[:, None] * y / fy[:, None]) / fx[:, None] * z
    y_lift = (y - cy[:, None]) / fy[:, None] * z
    pts_cam = torch.stack((x_lift, y_lift, z), dim=-1).permute(0, 2, 1)
    world = (torch.bmm(pose[:, :3, :3], pts_cam) + pose[:, :3, 3:]).permute(0, 2, 1)
    dirs = F.normalize(world - cam_loc[:, None, :], dim=2)
    return dirs, cam_loc


# ----------------------------------------------------------------------------- ragged glue
def ragged_query(grid: OracleGrid, pts: torch.Tensor, k: int, r: float, smax: int):
    """utils.py:90-113 on top of knnquery.py:168-285.  pts [R,D,3].
    Returns neighbor_idx [V,k] (int64, -1 pad), shading_pts [V,3], mask [R,smax] bool, ray_mask [R] bool."""
    out = (grid.query_dense_cdist if KNN_BACKEND == "cdist" else grid.query_dense)(pts, k, r, smax)
    ray_mask = out["ray_mask2"].bool()
    slot_valid = (out["pidx"] >= 0).any(-1)                       # [R,smax]
    mask = slot_valid & ray_mask[:, None]
    neighbor_idx = out["pidx"][mask].long()
    shading_pts = out["sample_loc"][mask]
    return neighbor_idx, shading_pts, mask, ray_mask


def _pairs(neighbor_idx):
    valid = neighbor_idx >= 0
    idx = torch.arange(neighbor_idx.shape[0])[:, None].expand_as(valid)[valid]   # utils.py:172-183
    nbr = neighbor_idx[valid]
    return valid, idx, nbr


def rbf_weights(x_pi, idx, V, rbf):
    """pointneus_disent.py:241-247 (weights carry no gradient)."""
    dist = torch.clamp(torch.norm(x_pi, dim=-1), min=1e-12).clone().detach()
    w = torch.exp(-((dist * rbf) ** 2))
    norm = torch.zeros(V).index_add_(0, idx, w)
    return w, norm


def aggregate_sdf(p: Params, x_pi, feat_g, w, norm, idx, V):
    """pointneus_disent.py:300-313"""
    h = mlp(torch.cat([feat_g, x_pi], -1), p.F_geometry)
    sdf = F.linear(h, *p.T)
    agg = torch.zeros(V, 1).index_add_(0, idx, w[:, None] * sdf)
    return agg / norm[:, None]


def point_sdf(p: Params, grid: OracleGrid, x: torch.Tensor, compact: bool = False):
    """sdf_importance / get_sdf_eval (fill 1000) and pseudo_sdf (compact=True: valid rows only, [V,1])
    (pointneus_disent.py:249-298, 348-421, 423-495)."""
    nidx, _, mask, ray_mask = ragged_query(grid, x[:, None, :], p.k, p.r, 1)
    filler = torch.ones(x.shape[0]) * 1000
    if nidx.shape[0] == 0:
        return filler
    valid, idx, nbr = _pairs(nidx)
    V = nidx.shape[0]
    sp = x[mask[:, 0]]
    x_pi = sp[idx] - p.neural_pts[nbr]
    w, norm = rbf_weights(x_pi, idx, V, p.rbf)
    agg = aggregate_sdf(p, x_pi, p.neural_feats_geometry[nbr], w, norm, idx, V)
    if compact:
        return agg
    filler = filler.clone()
    filler[ray_mask] = agg.squeeze(-1)
    return filler


# ----------------------------------------------------------------------------- sampler
@dataclass
class SamplerCfg:
    near: float = 0.5
    far: float = 6.0              # 2 * scene_bounding_sphere (ray_sampler.py:353)
    N_samples: int = 64
    N_samples_eval: int = 128
    N_samples_extra: int = 32
    eps: float = 0.1
    beta_iters: int = 10
    max_total_iters: int = 5
    add_tiny: float = 0.0


def error_bound(beta, sdf, dists, d_star):
    """ray_sampler.py:576-588 (beta scalar or [R,1])"""
    density = laplace_density(sdf, beta)
    shifted = torch.cat([torch.zeros(dists.shape[0], 1), dists * density[:, :-1]], dim=-1)
    integral = torch.cumsum(shifted, dim=-1)
    err_sec = torch.exp(-d_star / beta) * (dists ** 2.0) / (4 * beta ** 2)
    err_int = torch.cumsum(err_sec, dim=-1)
    bound = (torch.clamp(torch.exp(err_int), max=1.0e6) - 1.0) * torch.exp(-integral[:, :-1])
    return bound.max(-1)[0]


def sample_z(p: Params, grid: OracleGrid, ray_dirs, cam_loc, cfg: SamplerCfg, training: bool, fast: int = -1,
             rng: Optional[Dict[str, torch.Tensor]] = None, sdf_fn=None, trace: Optional[dict] = None,
             force_iters: Optional[int] = None):
    """ErrorBoundSampler_pn.get_z_vals (ray_sampler.py:377-574) + UniformSampler (:34-59).

    rng (training): {"t_rand":[R,N_eval], "u":[R,N_samples], "sampling_idx":[N_extra] long}.
    Returns z_vals [R, N_samples + N_extra + 2].

    force_iters (tests only): the outer loop is batch-global -- it runs while ANY ray of the batch is unconverged
    (ray_sampler.py:466-468) and every iteration resamples every ray -- so a SUBSET of a batch reproduces the batch's
    samples only if it runs the batch's number of iterations.  force_iters = that number replaces the subset's own
    `not_converge` decision; everything else is unchanged.
    """
    R = ray_dirs.shape[0]
    max_total_iters = fast if fast >= 0 else cfg.max_total_iters
    beta0 = get_beta(p).detach()
    sdf_fn = sdf_fn or (lambda pts: point_sdf(p, grid, pts))
    near = cfg.near * torch.ones(R, 1)
    far = cfg.far * torch.ones(R, 1)
    t_vals = torch.linspace(0.0, 1.0, steps=cfg.N_samples_eval)
    z_vals = near * (1.0 - t_vals) + far * t_vals
    if training:
        mids = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
        upper = torch.cat([mids, z_vals[..., -1:]], -1)
        lower = torch.cat([z_vals[..., :1], mids], -1)
        z_vals = lower + (upper - lower) * rng["t_rand"]
    samples, samples_idx = z_vals, None
    dists = z_vals[:, 1:] - z_vals[:, :-1]
    bound = (1.0 / (4.0 * torch.log(torch.tensor(cfg.eps + 1.0)))) * (dists ** 2.0).sum(-1)
    beta = torch.sqrt(bound)
    total_iters, not_converge = 0, True
    sdf = None
    while not_converge and total_iters < max_total_iters:
        points = cam_loc.unsqueeze(1) + samples.unsqueeze(2) * ray_dirs.unsqueeze(1)
        with torch.no_grad():
            samples_sdf = sdf_fn(points.reshape(-1, 3))
        if samples_idx is not None:
            merged = torch.cat([sdf.reshape(-1, z_vals.shape[1] - samples.shape[1]),
                                samples_sdf.reshape(-1, samples.shape[1])], -1)
            sdf = torch.gather(merged, 1, samples_idx).reshape(-1, 1)
        else:
            sdf = samples_sdf
        d = sdf.reshape(z_vals.shape)
        dists = z_vals[:, 1:] - z_vals[:, :-1]
        a, b, c = dists, d[:, :-1].abs(), d[:, 1:].abs()
        first = a.pow(2) + b.pow(2) <= c.pow(2)
        second = a.pow(2) + c.pow(2) <= b.pow(2)
        d_star = torch.zeros(z_vals.shape[0], z_vals.shape[1] - 1)
        d_star[first] = b[first]
        d_star[second] = c[second]
        s = (a + b + c) / 2.0
        area = s * (s - a) * (s - b) * (s - c)
        m = ~first & ~second & (b + c - a > 0)
        d_star[m] = (2.0 * torch.sqrt(area[m])) / (a[m])
        d_star = (d[:, 1:].sign() * d[:, :-1].sign() == 1) * d_star
        curr = error_bound(beta0, d, dists, d_star)
        beta[curr <= cfg.eps] = beta0
        beta_min, beta_max = beta0.unsqueeze(0).repeat(R), beta
        for _ in range(cfg.beta_iters):
            beta_mid = (beta_min + beta_max) / 2.0
            curr = error_bound(beta_mid.unsqueeze(-1), d, dists, d_star)
            beta_max[curr <= cfg.eps] = beta_mid[curr <= cfg.eps]
            beta_min[curr > cfg.eps] = beta_mid[curr > cfg.eps]
        beta = beta_max
        density = laplace_density(d, beta.unsqueeze(-1))
        dists = torch.cat([dists, torch.full((R, 1), 1e10)], -1)
        free_energy = dists * density
        shifted = torch.cat([torch.zeros(R, 1), free_energy[:, :-1]], dim=-1)
        alpha = 1 - torch.exp(-free_energy)
        transmittance = torch.exp(-torch.cumsum(shifted, dim=-1))
        weights = alpha * transmittance
        total_iters += 1
        not_converge = bool(beta.max() > beta0) if force_iters is None else total_iters < force_iters
        more = not_converge and total_iters < max_total_iters
        if trace is not None:
            trace.setdefault("iters", []).append(dict(z=z_vals.clone(), sdf=d.clone(), d_star=d_star.clone(),
                                                      beta=beta.clone(), weights=weights.clone()))
        bins = z_vals
        if more:
            N = cfg.N_samples_eval
            err_sec = torch.exp(-d_star / beta.unsqueeze(-1)) * (dists[:, :-1] ** 2.0) / (4 * beta.unsqueeze(-1) ** 2)
            err_int = torch.cumsum(err_sec, dim=-1)
            bound_opacity = (torch.clamp(torch.exp(err_int), max=1.0e6) - 1.0) * transmittance[:, :-1]
            pdf = bound_opacity + cfg.add_tiny
        else:
            N = cfg.N_samples
            pdf = weights[..., :-1] + 1e-5
        pdf = pdf / torch.sum(pdf, -1, keepdim=True)
        cdf = torch.cumsum(pdf, -1)
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
        if more or (not training):
            u = torch.linspace(0.0, 1.0, steps=N).unsqueeze(0).repeat(R, 1)
        else:
            u = rng["u"]
        u = u.contiguous()
        inds = torch.searchsorted(cdf, u, right=True)
        below = torch.max(torch.zeros_like(inds - 1), inds - 1)
        above = torch.min((cdf.shape[-1] - 1) * torch.ones_like(inds), inds)
        cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
        bin_b, bin_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
        denom = cdf_a - cdf_b
        denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
        t = (u - cdf_b) / denom
        samples = bin_b + t * (bin_a - bin_b)
        if more:
            z_vals, samples_idx = torch.sort(torch.cat([z_vals, samples], -1), -1)
    z_samples = samples
    if cfg.N_samples_extra > 0:
        if training:
            sampling_idx = rng["sampling_idx"]
        else:
            sampling_idx = torch.linspace(0, z_vals.shape[1] - 1, cfg.N_samples_extra).long()
        z_extra = torch.cat([near, far, z_vals[:, sampling_idx]], -1)
    else:
        z_extra = torch.cat([near, far], -1)
    z_out, _ = torch.sort(torch.cat([z_samples, z_extra], -1), -1)
    return z_out


# ----------------------------------------------------------------------------- TV regulariser
def tv_regul(p: Params, grid: OracleGrid):
    """utils.py:221-281"""
    kp_pos = p.neural_pts.detach()
    kp_feat = p.neural_feats_geometry
    n = kp_pos.shape[0]
    nidx, _, kmask, _ = ragged_query(grid, kp_pos[:, None, :], p.k, p.r, 1)
    padded = torch.full((n, p.k), -1, dtype=torch.long)
    padded[:, 0] = torch.arange(n)
    padded[kmask[:, 0]] = nidx
    origin = torch.arange(n)[:, None]
    ident = padded == origin
    enough = (padded >= 0).int().sum(-1, keepdim=True) > 1
    padded[ident & enough] = -1
    valid, idx, nbr = _pairs(padded)
    w = 1 / (torch.linalg.norm(kp_pos[nbr] - kp_pos[idx], dim=-1) + 1.0e-5)
    norm = torch.zeros(n).index_add_(0, idx, w)
    fd = torch.linalg.norm(kp_feat[nbr] - kp_feat[idx], ord=1, dim=-1)
    tv = torch.zeros(n).index_add_(0, idx, w * fd)
    return (tv / norm).mean()


# ----------------------------------------------------------------------------- full forward
def render_forward(p: Params, grid: OracleGrid, uv, pose, intrinsics, cfg: SamplerCfg, training: bool,
                   fast: int = -1, rng=None, far_cfg: float = 4.5, with_tv: bool = True, z_vals=None, local_data=None):
    """PointVolSDF.forward (pointneus_disent.py:614-892); the DTU-only local loss (:727-763) when `local_data` is given."""
    ray_dirs, cam_loc = camera_rays(uv, pose, intrinsics)
    ray_dirs_tmp, _ = camera_rays(uv, torch.eye(4)[None], intrinsics)
    depth_scale = ray_dirs_tmp[0, :, 2:]
    ray_dirs = ray_dirs.reshape(-1, 3)
    R = ray_dirs.shape[0]
    cam_loc = cam_loc.unsqueeze(1).repeat(1, R, 1).reshape(-1, 3)
    if z_vals is None:
        z_vals = sample_z(p, grid, ray_dirs, cam_loc, cfg, training, fast, rng)
    points = cam_loc.unsqueeze(1) + z_vals.unsqueeze(2) * ray_dirs.unsqueeze(1)
    S = p.max_shading_pts
    nidx, shading_pts, mask, ray_mask = ragged_query(grid, points, p.k, p.r, S)
    vm = mask[ray_mask]                                            # valid_neural_pts_mask [Rv,S]
    out = {"z_vals": z_vals, "ray_mask": ray_mask, "mask": mask}
    pseudo_loss = torch.tensor(0.0)
    local_loss = torch.tensor(0.0)
    Rv = int(ray_mask.sum())
    have = shading_pts.shape[0] > 0
    if have:
        # filter_points (pointneus_disent.py:207-239)
        o, dd = cam_loc[ray_mask], ray_dirs[ray_mask]
        sqp = torch.zeros(Rv, S, 3)
        sqp[vm] = shading_pts.clone().detach()
        t = ((sqp - o.unsqueeze(1)) / dd.unsqueeze(1)).nanmean(dim=-1, keepdim=True)
        z_values = torch.zeros_like(t)
        z_values[vm] = t[vm]
        _z = torch.cat([z_values, torch.zeros(Rv, 1, 1)], dim=1)
        deltas = _z[:, 1:] - _z[:, :-1]
        deltas[~vm] = 0
        deltas = deltas.clamp_(min=0)
        shading_pts = (o.unsqueeze(1) + z_values * dd.unsqueeze(1))[vm]
        shading_pts.requires_grad_(True)
        valid, idx, nbr = _pairs(nidx)
        V = nidx.shape[0]
        x_pi = shading_pts[idx] - p.neural_pts[nbr]
        w, norm = rbf_weights(x_pi, idx, V, p.rbf)
        agg_sdf = aggregate_sdf(p, x_pi, p.neural_feats_geometry[nbr], w, norm, idx, V)
        gradients = torch.autograd.grad(agg_sdf, shading_pts, torch.ones_like(agg_sdf), retain_graph=True,
                                        create_graph=True)[0]
        # get_color (pointneus_disent.py:325-346)
        fin = torch.cat([positional_encoding(x_pi, 6), p.neural_feats_color[nbr]], -1)
        dirs_v = dd.unsqueeze(1).expand(-1, S, -1)[vm]
        if BF16_OPERANDS:
            colors = _color_bf16_operands(p, fin, dirs_v, w, norm, idx, V)
        else:
            fcol = mlp(fin, p.F_color)
            agg_feat = torch.zeros(V, 256).index_add_(0, idx, w[:, None] * fcol) / norm[:, None]
            h = mlp(torch.cat([positional_encoding(dirs_v, 3), agg_feat], -1), p.R)
            colors = torch.sigmoid(h)
        sdf_filler = torch.ones(Rv, S, 1) * 1000
        sdf_filler[vm] = agg_sdf
        density_filler = torch.zeros(Rv, S, 1)
        density_filler[vm] = laplace_density(agg_sdf, get_beta(p))
        weights_values = volume_rendering(deltas[..., 0], density_filler[..., 0])
        if local_data is not None and training:                                       # pointneus_disent.py:727-763
            from .local_loss import local_loss_from_rays
            local_loss = local_loss_from_rays(sdf_filler[..., 0], z_values[..., 0], o, dd, local_data)
        dist_map = torch.sum(weights_values / (weights_values.sum(-1, keepdim=True) + 1e-10) * z_values.squeeze(-1), -1)
        pts_rendered = o + dd * dist_map[:, None]
        sdf_rendered = point_sdf(p, grid, pts_rendered, compact=True)
        pseudo_loss = F.l1_loss(sdf_rendered, torch.zeros_like(sdf_rendered), reduction="mean")
        out["pseudo_count"] = int(sdf_rendered.shape[0])   # the mean's denominator (data-parallel tests)
        color_filler = torch.zeros(Rv, S, 3)
        color_filler[vm] = colors
        rgb_values = torch.sum(weights_values.unsqueeze(-1) * color_filler, 1)
        depth_values = torch.sum(weights_values * z_values.squeeze(-1), 1, keepdim=True) / (
            weights_values.sum(dim=1, keepdim=True) + 1e-8)
        acc_values = torch.sum(weights_values, -1, keepdim=True)
        if not training:
            normals = torch.zeros(Rv, S, 3)
            g = gradients.detach()
            normals[vm] = g / g.norm(2, -1, keepdim=True)
            normal_values = torch.sum(weights_values.unsqueeze(-1) * normals, 1)
        points_filler = torch.zeros(Rv, S, 3)
        points_filler[vm] = shading_pts
        out.update(sdf=sdf_filler, deltas=deltas, z_values=z_values, colors=color_filler, dist_map=dist_map,
                   agg_sdf=agg_sdf, valid_mask=vm, neighbor_idx=nidx)
    rgb = torch.zeros(R, 3)
    normal = torch.zeros(R, 3)
    acc = torch.zeros(R, 1)
    depth = torch.full((R, 1), 1.0)
    weights = torch.zeros(R, S)
    depth_vals = torch.ones(R, S) * far_cfg
    xyz = torch.zeros(R, S, 3)
    if have:
        xyz[ray_mask] = points_filler
        rgb[ray_mask] = rgb_values
        acc[ray_mask] = acc_values
        depth[ray_mask] = depth_values
        weights[ray_mask] = weights_values
        depth_vals[ray_mask] = z_values.squeeze(-1) * depth_scale[ray_mask]
    out.update(rgb_values=rgb, depth_values=depth, depth_vals=depth_vals, weights=weights, xyz=xyz,
               accumulation=acc, local_loss=local_loss, pseudo_pts_loss=pseudo_loss)
    out["tv_loss"] = tv_regul(p, grid) if with_tv else torch.tensor(0.0)
    if not training:
        if have:
            normal[ray_mask] = normal_values
        out["normal_map"] = normal
    elif have:
        out["grad_theta"] = gradients
    return out


def volsdf_loss(out, rgb_gt, mask_gt, rgb_weight=1.0, eikonal_weight=0.001, tv_weight=0.01, local_weight=0.5,
                pseudo_weight=0.5):
    """loss.py:51-100 (weights: config/ours.yaml:15-20). rgb_gt [R,3], mask_gt [R,1] float."""
    res = {"rgb_loss": F.l1_loss(out["rgb_values"], rgb_gt.reshape(-1, 3))}
    if "grad_theta" in out:
        res["eikonal_loss"] = ((out["grad_theta"].norm(2, dim=1) - 1) ** 2).mean()
    else:
        res["eikonal_loss"] = torch.tensor(0.0)
    res["tv_loss"] = out["tv_loss"] if tv_weight > 0 else torch.tensor(0.0)
    wsum = out["weights"].sum(-1, keepdim=True)
    res["mask_loss"] = F.binary_cross_entropy(wsum.clip(1e-3, 1.0 - 1e-3), mask_gt.reshape(-1, 1))
    res["local_loss"] = out["local_loss"]
    res["pseudo_loss"] = out["pseudo_pts_loss"] if pseudo_weight > 0 else torch.tensor(0.0)
    res["loss"] = (rgb_weight * res["rgb_loss"] + eikonal_weight * res["eikonal_loss"] + tv_weight * res["tv_loss"]
                   + local_weight * res["local_loss"] + pseudo_weight * res["pseudo_loss"] + res["mask_loss"])
    return res
