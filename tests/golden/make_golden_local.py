"""Golden vectors for the feature-consistency ("local") loss, produced by the REFERENCE's own code on CPU:
``spurfies.feat_utils.get_local_loss`` (feat_utils.py:377-451) and ``PointVolSDF.find_surface_points``
(pointneus_disent.py:586-612), called exactly as pointneus_disent.py:727-763 calls them.

Runs only in the authoring container (needs /root/reference; same import shims as make_golden.py).  The feature
maps / cameras are NOT stored: tests regenerate them with ``spurfies_b200.scenes.local_data`` (checksums are stored).

Usage:  python tests/golden/make_golden_local.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

FEAT_RES = (128, 96)


def synthetic_rays(R=160, S=80, seed=3):
    """Dense [R,S] SDF / depth rows the way the model hands them over: ascending depths on valid slots, 1000 / 0 on
    slots without neighbours, SDF of a bumpy sphere plus noise (several sign changes on some rays, none on others)."""
    from spurfies_b200 import scenes
    g = torch.Generator().manual_seed(seed)
    cam = scenes.camera(0, 2.3)
    uv = (scenes.pixel_batch(R, seed=seed) - torch.tensor([256.0, 192.0])) * 0.5 + torch.tensor([256.0, 192.0])
    from oracle import hotpath as H
    dirs, o = H.camera_rays(uv, cam["pose"], cam["intrinsics"])
    dirs = dirs.reshape(-1, 3)
    o = o.reshape(1, 3).expand(R, 3).contiguous()
    z = torch.sort(torch.rand(R, S, generator=g) * 1.6 + 1.5, dim=1)[0]
    p = o[:, None] + z[..., None] * dirs[:, None]
    sdf = p.norm(dim=-1) - 0.42 + 0.03 * torch.sin(9 * p[..., 0]) + 0.01 * torch.randn(R, S, generator=g)
    invalid = torch.rand(R, S, generator=g) < 0.25
    invalid[: R // 8] = True                      # rays with no shading point at all
    invalid[R // 8: R // 6, 1:] = True            # rays with a single valid slot
    sdf = torch.where(invalid, torch.full_like(sdf, 1000.0), sdf)
    z = torch.where(invalid, torch.zeros_like(z), z)
    return sdf, z, o, dirs


def main():
    from make_golden import install_shims
    install_shims()
    from spurfies_b200 import scenes
    from spurfies.model import pointneus_disent as pd
    from spurfies import feat_utils

    sdf, z, o, dirs = synthetic_rays()
    ld = scenes.local_data(0, 2.3, feat_res=FEAT_RES)
    gold = {"sdf": sdf, "z": z, "cam_loc": o, "ray_dirs": dirs, "feat_res": FEAT_RES,
            "local_data_checksum": {k: float(v.double().abs().sum()) for k, v in ld.items()}}
    leaf = sdf.clone().requires_grad_(True)
    sdf_f = leaf * 1.0                               # the reference writes NaN in place (pointneus_disent.py:587)
    d_surface, network_mask = pd.PointVolSDF.find_surface_points(None, sdf_f.unsqueeze(0), z.unsqueeze(0), device="cpu")
    d_surface, network_mask = d_surface.squeeze(0), network_mask.squeeze(0)
    object_mask = network_mask
    point_surface = o + dirs * d_surface[:, None]
    pts = point_surface[network_mask & object_mask]
    size, center = ld["size"].unsqueeze(0)[:1], ld["center"].unsqueeze(0)[:1]
    loss = feat_utils.get_local_loss(pts, None, ld["feat"].unsqueeze(0), ld["cam"].unsqueeze(0),
                                     ld["feat_src"].unsqueeze(0), ld["src_cams"].unsqueeze(0), size, center,
                                     network_mask.reshape(-1), object_mask.reshape(-1))
    loss.backward()
    gold.update(d_surface=d_surface.detach(), network_mask=network_mask, surface_points=pts.detach(),
                loss=loss.detach(), d_sdf=torch.nan_to_num(leaf.grad, nan=0.0))
    # a second case: nothing crosses -> 0 (feat_utils.py:390-391)
    none = torch.full((4, 80), 1000.0)
    d2, m2 = pd.PointVolSDF.find_surface_points(None, none.clone().unsqueeze(0), torch.zeros(1, 4, 80), device="cpu")
    gold["empty_mask_sum"] = int(m2.sum())
    path = os.path.join(ROOT, "tests", "golden", "local_loss.pt")
    torch.save(gold, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1e3:.1f} kB): loss {float(loss):.6f}, hits {int(network_mask.sum())}/"
          f"{len(network_mask)}, |d_sdf| {float(gold['d_sdf'].abs().sum()):.4e}")


if __name__ == "__main__":
    main()
