"""GPU: the fused VolSDFLoss kernels (spf_volsdf_loss) against the torch restatement of loss.py:51-100 that
VolSDFLoss keeps for the reference's ragged contract -- values of every term and the gradients they send back."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R,S,with_eik", [(4096, 80, True), (37, 80, True), (513, 24, False)])
def test_fused_loss_matches_torch_path(R, S, with_eik):
    from spurfies_b200.model import VolSDFLoss
    g = torch.Generator().manual_seed(R)
    dev = "cuda"
    rgb = torch.rand(R, 3, generator=g).to(dev)
    w = (torch.rand(R, S, generator=g) * (2.2 / S)).to(dev)
    w[: R // 8] = 0.0                      # rays that miss: sum below the clip
    w[R // 8: R // 6] *= 3.0               # and some above it
    gt = {"rgb": torch.rand(1, R, 3, generator=g).to(dev), "mask": (torch.rand(1, R, 3, generator=g) > 0.3).float().to(dev)}
    grad = torch.randn(R * S, 3, generator=g).to(dev)
    mask = (torch.rand(R * S, generator=g) > 0.5).to(dev)
    res = {}
    for mode in ("torch", "fused"):
        leaves = {"rgb": rgb.clone().requires_grad_(), "w": w.clone().requires_grad_(),
                  "tv": torch.tensor(0.37, device=dev, requires_grad=True),
                  "ps": torch.tensor(0.11, device=dev, requires_grad=True),
                  "lo": torch.tensor(0.05, device=dev, requires_grad=True)}
        out = {"rgb_values": leaves["rgb"], "weights": leaves["w"], "tv_loss": leaves["tv"], "pseudo_pts_loss": leaves["ps"],
               "local_loss": leaves["lo"]}
        if with_eik:
            out["grad_theta_dense"], out["grad_theta_mask"] = grad, mask
        loss_mod = VolSDFLoss()
        lo = _torch_path(loss_mod, out, gt) if mode == "torch" else loss_mod(out, gt)
        lo["loss"].backward()
        res[mode] = ({k: float(v) for k, v in lo.items()}, {k: v.grad.clone() for k, v in leaves.items()})
    for k, v in res["torch"][0].items():
        assert abs(res["fused"][0][k] - v) <= 1e-5 * max(1.0, abs(v)), (k, res["fused"][0][k], v)
    for k, v in res["torch"][1].items():
        a = res["fused"][1][k]
        # d weights: BCE'(x) = (x - y) / (x (1 - x)); for sums near the upper clip 1 - x cancels catastrophically, so the
        # two summation orders of sum_s w differ by ulp(1) / (1 - x) ~ 6e-5 relative there
        tol = 3e-4 if k == "w" else 1e-5
        assert float((a - v).abs().max()) <= 1e-6 + tol * float(v.abs().max()), k


def _torch_path(loss_mod, out, gt):
    """loss.py:51-100 in plain torch ops on the dense (sample, mask) form of grad_theta."""
    import torch.nn.functional as F
    dev = out["rgb_values"].device
    rgb_gt = gt["rgb"].reshape(-1, 3)
    o = {"rgb_loss": (out["rgb_values"] - rgb_gt).abs().mean()}
    if "grad_theta_dense" in out:
        g, m = out["grad_theta_dense"], out["grad_theta_mask"]
        o["eikonal_loss"] = (((g.norm(2, dim=1) - 1) ** 2) * m).sum() / m.sum().clamp(min=1)
    else:
        o["eikonal_loss"] = torch.zeros((), device=dev)
    wsum = out["weights"].sum(-1, keepdim=True)
    o["mask_loss"] = F.binary_cross_entropy(wsum.clip(1e-3, 1.0 - 1e-3), gt["mask"].squeeze()[:, 0][..., None])
    o["tv_loss"], o["local_loss"], o["pseudo_loss"] = out["tv_loss"], out["local_loss"], out["pseudo_pts_loss"]
    o["loss"] = (loss_mod.rgb_weight * o["rgb_loss"] + loss_mod.eikonal_weight * o["eikonal_loss"] + loss_mod.tv_weight * o["tv_loss"]
                 + loss_mod.local_weight * o["local_loss"] + loss_mod.pseudo_weight * o["pseudo_loss"] + o["mask_loss"])
    return o


@pytest.mark.parametrize("world", [2, 8, 3])
def test_tv_point_slices_average_to_the_whole_regulariser(world):
    """spf_tv_fwd_bwd_range: a data-parallel rank evaluates 1/world of the points scaled by world; the AVERAGE over the
    ranks (what the gradient all-reduce + 1/world computes) is the whole tv_regul (utils.py:221-281)."""
    from spurfies_b200 import scenes
    from spurfies_b200.dist import shard_range
    from spurfies_b200.fields import TVRegul
    from spurfies_b200.model import PointVolSDF, default_conf
    sc = scenes.dtu_like(6000, seed=5, radii=(0.35, 0.5))
    model = PointVolSDF(default_conf(), "24", "dtu", neural_points=sc["pts"], neural_colors=sc["colors"]).cuda()
    pts = model.neural_pts
    nbr = model._grid().query_points(pts, model.conf.k, model.conf.r)
    feat = torch.randn(pts.shape[0], 32, generator=torch.Generator().manual_seed(1)).cuda().requires_grad_()
    whole = TVRegul.apply(feat, pts, nbr)
    whole.backward()
    g_whole, feat.grad = feat.grad.clone(), None
    val, g = 0.0, torch.zeros_like(g_whole)
    for r in range(world):
        lo, hi = shard_range(pts.shape[0], r, world)
        v = TVRegul.apply(feat, pts, nbr, lo, hi - lo, float(world))
        v.backward()
        val += float(v) / world
        g += feat.grad / world
        feat.grad = None
    assert abs(val - float(whole)) <= 1e-5 * abs(float(whole))
    assert float((g - g_whole).abs().max()) <= 1e-5 * float(g_whole.abs().max())
