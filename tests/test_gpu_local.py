"""GPU: SURVEY 8(f4) -- the feature-consistency ("local") loss kernels (spf_local_loss_fwd / _bwd) against golden vectors
produced by the reference's own feat_utils.get_local_loss + find_surface_points (tests/golden/make_golden_local.py), and
the full training forward with `local_data` against the oracle.  fp32, 1e-4 relative."""
import os

import pytest
import torch

from oracle import hotpath as H
from oracle import local_loss as OL
from tests.helpers import load_golden, load_into_model, rel_err, trainable

pytestmark = pytest.mark.gpu
TOL = 1e-4
GOLD = os.path.join(os.path.dirname(__file__), "golden", "local_loss.pt")


def _cuda(d):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()}


@pytest.fixture(scope="module")
def gold():
    from spurfies_b200 import scenes
    g = torch.load(GOLD)
    return g, scenes.local_data(0, 2.3, feat_res=tuple(g["feat_res"]))


def _run(g, ld, sdf=None):
    from spurfies_b200.fields import LocalLoss, local_feature_args
    sdf = (g["sdf"] if sdf is None else sdf).cuda().reshape(-1).requires_grad_(True)
    R, S = g["sdf"].shape
    feats = local_feature_args(ld, sdf.device)
    loss, d_surface, cross = LocalLoss.apply(sdf, g["z"].cuda().contiguous(), g["cam_loc"][0].cuda().contiguous(),
                                             g["ray_dirs"].cuda().contiguous(), feats, R, S)
    loss.backward()
    return loss.detach(), d_surface, cross, sdf.grad.view(R, S)


@pytest.mark.parametrize("layout", ["nchw", "channels_last"])
def test_kernel_matches_reference_golden(gold, layout):
    from spurfies_b200.fields import channels_last_features
    g, ld = gold
    ld = _cuda(ld)
    if layout == "channels_last":
        ld = channels_last_features(ld)
        assert ld["feat"].stride(0) == 1 and ld["feat"].shape == gold[1]["feat"].shape
    loss, d_surface, cross, d_sdf = _run(g, ld)
    assert torch.equal((cross >= 0).cpu(), g["network_mask"])
    assert rel_err(d_surface, g["d_surface"]) < 1e-6
    assert abs(float(loss) - float(g["loss"])) < TOL * float(g["loss"])
    assert rel_err(d_sdf, g["d_sdf"]) < TOL
    assert not bool((d_sdf != 0).cpu()[~g["network_mask"]].any())


def test_find_surface_points_method(gold):
    from spurfies_b200.model import PointVolSDF, default_conf
    g, _ = gold
    gg, P = load_golden()
    model = PointVolSDF(default_conf(), "24", "dtu", neural_points=gg["scene"]["pts"], neural_colors=gg["scene"]["colors"])
    d, m = model.find_surface_points(g["sdf"].cuda().unsqueeze(0), g["z"].cuda().unsqueeze(0))
    assert d.shape == (1, 160) and m.shape == (1, 160)
    assert torch.equal(m[0].cpu(), g["network_mask"]) and rel_err(d[0], g["d_surface"]) < 1e-6


def test_edge_cases(gold):
    g, ld = gold
    ld = _cuda(ld)
    # no slot has a neighbour / no sign change: loss 0, zero gradient (feat_utils.py:390-391)
    for sdf in (torch.full_like(g["sdf"], 1000.0), g["sdf"].abs() + 0.1):
        loss, d_surface, cross, d_sdf = _run(g, ld, sdf)
        assert float(loss) == 0.0 and int((cross >= 0).sum()) == 0 and float(d_sdf.abs().sum()) == 0.0
        assert float(d_surface.abs().sum()) == 0.0
    # one source view (m = 1) against the oracle restatement
    one = dict(ld, feat_src=ld["feat_src"][:1], src_cams=ld["src_cams"][:1])
    loss, _, _, d_sdf = _run(g, one)
    s = g["sdf"].clone().requires_grad_(True)
    want = OL.local_loss_from_rays(s, g["z"], g["cam_loc"], g["ray_dirs"], {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in one.items()})
    want.backward()
    assert abs(float(loss) - float(want)) < TOL * float(want)
    assert rel_err(d_sdf, torch.nan_to_num(s.grad)) < TOL
    # identical source and reference views: cosine 1 everywhere -> loss ~ 0
    same = dict(ld, feat_src=ld["feat"][None].repeat(2, 1, 1, 1), src_cams=ld["cam"][None].repeat(2, 1, 1, 1))
    loss, _, _, d_sdf = _run(g, same)
    assert float(loss) < 1e-6


def test_training_forward_with_local_data_matches_oracle():
    """Whole training step with the DTU feature term switched on: model (CUDA) vs oracle, loss and gradients."""
    from spurfies_b200 import scenes
    from spurfies_b200.model import PointVolSDF, VolSDFLoss, default_conf
    g, P = load_golden()
    # the random-init prior of the fixture is ~ -0.06 everywhere: centre it so that the SDF changes sign along the rays
    P.T = (P.T[0], P.T[1] + 0.0594)
    model = load_into_model(PointVolSDF(default_conf(), "24", "dtu", neural_points=g["scene"]["pts"],
                                        neural_colors=g["scene"]["colors"]), P)
    model.train()
    R = 96
    cam = scenes.camera(0, 2.3)
    uv = (scenes.pixel_batch(R, seed=9) - torch.tensor([256.0, 192.0])) * 0.45 + torch.tensor([256.0, 192.0])
    rng, gt = scenes.rng_inputs(R, step=2), scenes.synthetic_gt(R, 2)
    ld = scenes.local_data(0, 2.3, feat_res=(128, 96), size=2.0, center=(0.0, 0.0, 0.0))
    out = model({"intrinsics": cam["intrinsics"].cuda(), "uv": uv.cuda(), "pose": cam["pose"].cuda(), "local_data": _cuda(ld)},
                fast=1, rng=_cuda(rng))
    lo = VolSDFLoss()(out, _cuda(gt))
    lo["loss"].backward()
    Pt = trainable(P)
    ro = H.render_forward(Pt, Pt.make_grid(), uv, cam["pose"], cam["intrinsics"], H.SamplerCfg(), True, 1, rng, local_data=ld)
    rl = H.volsdf_loss(ro, gt["rgb"], gt["mask"][0, :, 0])
    rl["loss"].backward()
    assert float(ro["local_loss"]) > 1e-3, "the synthetic step must exercise the feature term"
    assert abs(float(out["local_loss"]) - float(ro["local_loss"])) < TOL * float(ro["local_loss"])
    assert abs(float(lo["loss"]) - float(rl["loss"])) < TOL * float(rl["loss"])
    assert rel_err(model.neural_feats_geometry.grad, Pt.neural_feats_geometry.grad) < 1e-3
    # eval mode never computes it (pointneus_disent.py:727 `and self.training`)
    model.eval()
    with torch.no_grad():
        oe = model({"intrinsics": cam["intrinsics"].cuda(), "uv": uv.cuda(), "pose": cam["pose"].cuda(),
                    "local_data": _cuda(ld)}, fast=1)
    assert float(oe["local_loss"]) == 0.0


def test_full_size_properties():
    """BASELINE configs[1] sizes (4096 rays x 80 slots, 384x512 feature maps): size-independent properties."""
    from spurfies_b200 import scenes
    from spurfies_b200.fields import LocalLoss, local_feature_args
    R, S = 4096, 80
    gen = torch.Generator().manual_seed(0)
    cam = scenes.camera(1, 2.3)
    dirs, o = H.camera_rays(scenes.pixel_batch(R, seed=3), cam["pose"], cam["intrinsics"])
    dirs, o = dirs.reshape(-1, 3).cuda(), o.reshape(3).cuda()
    z = torch.sort(torch.rand(R, S, generator=gen) * 1.6 + 1.5, dim=1)[0].cuda()
    p = o + z[..., None] * dirs[:, None]
    sdf = (p.norm(dim=-1) - 0.45).reshape(-1).requires_grad_(True)
    ld = _cuda(scenes.local_data(1, 2.3, feat_res=(512, 384), size=2.0, center=(0.0, 0.0, 0.0)))
    feats = local_feature_args(ld, "cuda")
    loss, d_surface, cross = LocalLoss.apply(sdf, z, o, dirs, feats, R, S)
    loss.backward()
    hit = cross >= 0
    assert 100 < int(hit.sum()) < R
    assert 0.0 <= float(loss) <= 0.5                               # every kept term is < 0.5
    # crossing depth lies between the two bracketing samples and on the sphere
    c = cross[hit].long()
    zz = z[hit]
    lo_, hi_ = zz.gather(1, c[:, None])[:, 0], zz.gather(1, c[:, None] + 1)[:, 0]
    ds = d_surface[hit]
    assert bool(((ds >= lo_) & (ds <= hi_)).all())
    # the interpolated crossing of the exact SDF of a sphere lies on it up to the chord error (sample spacing ~ 0.02)
    assert float(((o + ds[:, None] * dirs[hit]).norm(dim=-1) - 0.45).abs().max()) < 5e-3
    g = sdf.grad.view(R, S)
    nz = g != 0
    assert int(nz.sum()) <= 2 * int(hit.sum()) and not bool(nz[~hit].any())
    assert bool(torch.isfinite(g).all())
