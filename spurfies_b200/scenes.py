"""Synthetic DTU- / Mip-NeRF-360-shaped scenes, cameras and injected RNG draws.

There is no network for datasets or checkpoints, so every test and benchmark runs on these
(SURVEY.md section 8(d)).  Shapes and constants follow the reference configuration:
grid (0.025 x3 voxels, kernel 3) `spurfies/model/pointneus_disent.py:45-62`; cameras are pinhole
4x4 intrinsics / cam-to-world poses as consumed by `spurfies/utils/rend_util.py:60-95`;
RNG draws replace the CPU `torch.rand` / `randperm` calls of `spurfies/model/ray_sampler.py:55, 514, 550`.
Host-side numpy/torch only; nothing here touches the GPU.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _bumpy_sphere(rng, n, radius, bump=0.04, jitter=0.002):
    d = _unit(rng.normal(size=(n, 3)))
    r = radius * (1.0 + bump * np.sin(7 * d[:, 0]) * np.cos(5 * d[:, 1]) + bump * 0.5 * np.sin(9 * d[:, 2]))
    return d * r[:, None] + rng.normal(scale=jitter, size=(n, 3))


def dtu_like(n_points: int = 100_000, seed: int = 24, radii=(0.35, 0.5, 0.65)) -> Dict[str, torch.Tensor]:
    """Nested bumpy closed surfaces inside [-1,1]^3 (DTU scan-shaped object)."""
    rng = np.random.default_rng(seed)
    area = np.array([r * r for r in radii])
    counts = np.floor(n_points * area / area.sum()).astype(int)
    counts[-1] = n_points - counts[:-1].sum()
    pts = np.concatenate([_bumpy_sphere(rng, c, r) for c, r in zip(counts, radii)], 0)
    pts = np.clip(pts, -0.98, 0.98).astype(np.float32)
    rng.shuffle(pts, axis=0)
    colors = rng.integers(0, 256, size=(n_points, 3)).astype(np.float32)
    return {"pts": torch.from_numpy(pts), "colors": torch.from_numpy(colors), "ranges": (-1, -1, -1, 1, 1, 1),
            "cam_radius": 2.3, "name": "dtu_like"}


def garden_like(n_points: int = 1_000_000, seed: int = 360) -> Dict[str, torch.Tensor]:
    """Ground plane + central object + sparse background shell inside [-2,2]^3 (Mip-NeRF 360 garden-shaped)."""
    rng = np.random.default_rng(seed)
    n_ground = int(0.55 * n_points)
    n_obj = int(0.30 * n_points)
    n_bg = n_points - n_ground - n_obj
    g = np.stack([rng.uniform(-1.95, 1.95, n_ground), rng.uniform(-1.95, 1.95, n_ground),
                  -0.5 + 0.02 * np.sin(rng.uniform(0, 20, n_ground))], -1)
    g += rng.normal(scale=0.002, size=g.shape)
    obj = _bumpy_sphere(rng, n_obj, 0.45, bump=0.08)
    d = _unit(rng.normal(size=(n_bg, 3)))
    d[:, 2] = np.abs(d[:, 2]) * 0.6
    bg = _unit(d) * rng.uniform(1.6, 1.9, (n_bg, 1))
    pts = np.clip(np.concatenate([g, obj, bg], 0), -1.98, 1.98).astype(np.float32)
    rng.shuffle(pts, axis=0)
    colors = rng.integers(0, 256, size=(n_points, 3)).astype(np.float32)
    return {"pts": torch.from_numpy(pts), "colors": torch.from_numpy(colors), "ranges": (-2, -2, -2, 2, 2, 2),
            "cam_radius": 3.0, "name": "garden_like"}


def look_at_pose(eye: np.ndarray, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """cam-to-world 4x4, camera looks along +z (the convention `lift` + `pose[:3,:3] @ p_cam` implies)."""
    eye = np.asarray(eye, dtype=np.float64)
    f = _unit(np.asarray(target, dtype=np.float64) - eye)
    r = _unit(np.cross(f, np.asarray(up, dtype=np.float64)))
    d = np.cross(f, r)
    pose = np.eye(4)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = r, d, f, eye
    return pose.astype(np.float32)


def camera(view: int, radius: float, res: Tuple[int, int] = (512, 384), focal: float = 600.0, n_views: int = 3):
    """One of `n_views` pinhole cameras on an orbit of `radius`, looking at the origin. res = (W, H)."""
    ang = 2 * math.pi * view / n_views + 0.3
    eye = np.array([radius * math.cos(ang) * 0.94, radius * math.sin(ang) * 0.94, radius * 0.34])
    W, H = res
    K = np.eye(4, dtype=np.float32)
    K[0, 0] = K[1, 1] = focal
    K[0, 2], K[1, 2] = W / 2.0, H / 2.0
    return {"pose": torch.from_numpy(look_at_pose(eye))[None], "intrinsics": torch.from_numpy(K)[None], "res": res}


def pixel_batch(n_rays: int, seed: int, res: Tuple[int, int] = (512, 384)) -> torch.Tensor:
    """Random pixel subset, uv [1,R,2] float (spurfies/datasets/dtu.py:360-364 picks a random subset per step)."""
    g = torch.Generator().manual_seed(seed)
    W, H = res
    idx = torch.randperm(W * H, generator=g)[:n_rays] if n_rays <= W * H else torch.randint(W * H, (n_rays,), generator=g)
    uv = torch.stack([(idx % W).float(), (idx // W).float()], -1)
    return uv[None]


def full_image_uv(res: Tuple[int, int] = (512, 384)) -> torch.Tensor:
    W, H = res
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    return torch.stack([xs.reshape(-1).float(), ys.reshape(-1).float()], -1)[None]


def rng_inputs(n_rays: int, step: int, n_coarse: int = 128, n_fine: int = 64, n_extra: int = 32):
    """Injected RNG draws for one training step (CPU generator, like the reference's CPU torch.rand)."""
    g = torch.Generator().manual_seed(1_000_003 * step + 17)
    return {
        "t_rand": torch.rand(n_rays, n_coarse, generator=g),
        "u": torch.rand(n_rays, n_fine, generator=g),
        "sampling_idx": torch.randperm(n_coarse, generator=g)[:n_extra],
    }


def synthetic_gt(n_rays: int, seed: int):
    g = torch.Generator().manual_seed(seed + 99)
    return {"rgb": torch.rand(1, n_rays, 3, generator=g), "mask": (torch.rand(1, n_rays, 3, generator=g) > 0.3).float()}


def local_data(view: int, radius: float, n_views: int = 3, feat_res: Tuple[int, int] = (512, 384), channels: int = 32,
               size: float = 2.6, center=(0.05, -0.1, 0.15), seed: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic stand-in for the Vis-MVSNet feature data of one DTU training view (spurfies/datasets/dtu.py:205-240,
    277-291): `feat` [C,H,W] of the view, `feat_src` [m,C,H,W] of the other views, `cam` [2,4,4] / `src_cams` [m,2,4,4]
    (row 0 = world->camera of the UN-normalised scene, row 1[:3,:3] = intrinsics at twice the feature resolution),
    `size`, `center` of the normalisation (world = p / 2 * size + center).  Features are smooth positive functions of
    the pixel so that neighbouring views correlate (no network / checkpoint is available)."""
    W, H = feat_res
    g = torch.Generator().manual_seed(seed + 4242)
    freq = torch.rand(channels, 2, generator=g) * 5.0 + 0.5
    phase = torch.rand(channels, generator=g) * 6.2831853
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32) / H, torch.arange(W, dtype=torch.float32) / W,
                            indexing="ij")
    c = torch.tensor(center, dtype=torch.float32)
    to_norm = torch.eye(4)
    to_norm[:3, :3] *= 2.0 / size
    to_norm[:3, 3] = -2.0 * c / size

    def one(v):
        cam = camera(v, radius, res=(2 * W, 2 * H), focal=600.0 * (2 * W) / 512.0, n_views=n_views)
        ext = torch.linalg.inv(cam["pose"][0]) @ to_norm                     # world (un-normalised) -> camera
        f = 1.0 + 0.6 * torch.sin(freq[:, 0, None, None] * xs[None] * 6.2831853 + freq[:, 1, None, None] * ys[None] * 6.2831853
                                  + phase[:, None, None] + 0.35 * v)
        return f.contiguous(), torch.stack([ext, cam["intrinsics"][0]], 0)
    src = [v for v in range(n_views) if v != view]
    f0, c0 = one(view)
    fs, cs = zip(*[one(v) for v in src])
    return {"feat": f0, "feat_src": torch.stack(fs, 0), "cam": c0, "src_cams": torch.stack(cs, 0),
            "size": torch.tensor(size), "center": c}
