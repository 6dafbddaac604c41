"""TEST INFRASTRUCTURE ONLY (never imported by spurfies_b200/): CPU restatement of the DTU feature-consistency
("local") loss of the reference, SURVEY 8(f4):

  * ``find_surface_points``   <- spurfies/model/pointneus_disent.py:586-612 (first back-facing SDF zero crossing per ray)
  * ``project``               <- spurfies/feat_utils.py:43-55 (idx_world2cam, idx_cam2img)
  * ``local_loss``            <- spurfies/feat_utils.py:377-451 (get_local_loss, uncerts=None) as it is called from
                                 pointneus_disent.py:727-763: ONE reference view, m source views

Pinned: ``tests/golden/local_loss.pt`` holds the outputs and gradients of the reference's own
``feat_utils.get_local_loss`` / ``PointVolSDF.find_surface_points`` imported from /root/reference on CPU
(``tests/golden/make_golden_local.py``); ``tests/test_oracle_local.py`` checks this restatement against them.
"""
import torch
import torch.nn.functional as F


def find_surface_points(sdf: torch.Tensor, d_all: torch.Tensor):
    """sdf, d_all [Rv,S] (sdf == 1000 marks slots without neighbours) -> d_surface [Rv], network_mask [Rv].
    pointneus_disent.py:586-612.  The reference takes `torch.max` over a bool row for the crossing index; that is the
    FIRST crossing (torch returns the first maximal element)."""
    sdf = torch.where(sdf == 1000, torch.full_like(sdf, float("nan")), sdf)          # :587
    crossing = (sdf[:, 1:] * sdf[:, :-1] < 0) & (sdf[:, 1:] < sdf[:, :-1])           # :589-593
    network_mask = crossing.any(dim=1)
    first = torch.argmax(crossing.to(torch.int8), dim=1)                              # :595
    rows = network_mask.nonzero()[:, 0]                                               # only the rays that cross
    col = first[rows]
    s0, s1 = sdf[rows, col], sdf[rows, col + 1]                                       # :598-603
    d0, d1 = d_all[rows, col], d_all[rows, col + 1]                                   # :604-609
    d_surface = torch.zeros(sdf.shape[0], dtype=sdf.dtype)                            # :597
    d_surface = d_surface.index_put((rows,), (s0 * d1 - s1 * d0) / (s0 - s1))         # :610
    return d_surface, network_mask


def project(pts_world: torch.Tensor, cam_pack: torch.Tensor) -> torch.Tensor:
    """pts_world [n,3], cam_pack [v,2,4,4] (row 0 = world->camera 4x4, row 1[:3,:3] = intrinsics) -> pixels [v,n,2].
    feat_utils.py:43-55, including both homogeneous normalisations and the three `+ 1e-9`."""
    n = pts_world.shape[0]
    homo = torch.cat([pts_world, torch.ones(n, 1)], dim=1)                            # feat_utils.py:405-407
    cam = torch.einsum("vij,nj->vni", cam_pack[:, 0], homo)                           # :45
    cam = cam / (cam[..., 3:4] + 1e-9)                                                # :46
    xyz = cam[..., :3] / (cam[..., 3:4] + 1e-9)                                       # :52
    img = torch.einsum("vij,vnj->vni", cam_pack[:, 1, :3, :3], xyz)                   # :53
    img = img / (img[..., 2:3] + 1e-9)                                                # :54
    return img[..., :2]


def local_loss(points: torch.Tensor, feat: torch.Tensor, cam: torch.Tensor, feat_src: torch.Tensor,
               src_cams: torch.Tensor, size: torch.Tensor, center: torch.Tensor) -> torch.Tensor:
    """points [n,3] surface points (normalised scene), feat [C,H,W], cam [2,4,4], feat_src [m,C,H,W],
    src_cams [m,2,4,4], size scalar, center [3] -> scalar.  feat_utils.py:377-451 with one reference view."""
    n = points.shape[0]
    if n == 0:                                                                        # :390-391
        return torch.tensor(0.0)
    pts_world = points / 2 * size.reshape(1, 1) + center.reshape(1, 3)                # :402-404
    cam_pack = torch.cat([cam[None], src_cams], dim=0)                                # :408
    grid = project(pts_world, cam_pack)                                               # :409-412  [1+m,n,2]
    feat_pack = torch.cat([feat[None], feat_src], dim=0)                              # :414
    H, W = feat_pack.shape[2:]
    grid_n = ((grid / 2) / torch.tensor([W, H], dtype=grid.dtype) * 2 - 1).clamp(-1.1, 1.1)   # :415, 58-68
    in_range = ((grid_n <= 1) & (grid_n >= -1)).all(dim=-1).to(grid.dtype)            # :416, 71-77  [1+m,n]
    valid = (in_range[:1] * in_range[1:]) > 0.5                                       # :417-419     [m,n]
    g = F.grid_sample(feat_pack, grid_n[:, :, None, :], mode="bilinear", padding_mode="zeros",
                      align_corners=False)[..., 0]                                    # :420-426  [1+m,C,n]
    norm = g.norm(dim=1)                                                              # :429
    corr = (g[:1] * g[1:]).sum(dim=1) / norm[:1].clamp(min=1e-9) / norm[1:].clamp(min=1e-9)   # :430-434
    corr_loss = (1 - corr).abs()                                                      # :435
    diff = corr_loss < 0.5                                                            # :437
    return (corr_loss * valid * diff).mean()                                          # :438 (mean over m*n)


def local_loss_from_rays(sdf: torch.Tensor, z_values: torch.Tensor, cam_loc: torch.Tensor, ray_dirs: torch.Tensor,
                         local_data: dict) -> torch.Tensor:
    """pointneus_disent.py:727-763: sdf/z_values [Rv,S] of the rays that have shading points, cam_loc/ray_dirs [Rv,3]."""
    d_surface, network_mask = find_surface_points(sdf, z_values)
    point_surface = cam_loc + ray_dirs * d_surface[:, None]                           # :745-748
    return local_loss(point_surface[network_mask], local_data["feat"], local_data["cam"], local_data["feat_src"],
                      local_data["src_cams"], local_data["size"], local_data["center"])
