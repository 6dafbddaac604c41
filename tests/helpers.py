"""Shared test helpers: golden fixtures (generated from the reference's own Python modules by
tests/golden/make_golden.py) and the oracle <-> product parameter plumbing."""
import os

import torch

from oracle import hotpath as H

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_dtu4k.pt")


def load_golden():
    g = torch.load(GOLDEN, weights_only=False)
    rec = g["params_recipe"]
    P = H.init_params(g["scene"]["pts"], g["scene"]["colors"], seed=rec["seed"])
    P.neural_feats_geometry *= rec["geometry_latent_scale"]
    P.neural_feats_color[:, 3:] *= rec["color_latent_scale_from3"]
    chk = {"neural_feats_color": P.neural_feats_color, "neural_feats_geometry": P.neural_feats_geometry,
           "T.w": P.T[0], "beta": P.beta}
    chk.update({f"F_color.{i}": W for i, (W, b) in enumerate(P.F_color)})
    chk.update({f"F_geometry.{i}": W for i, (W, b) in enumerate(P.F_geometry)})
    chk.update({f"R.{i}": W for i, (W, b) in enumerate(P.R)})
    for n, t in chk.items():
        got, want = float(t.double().abs().sum()), g["params_checksum"][n]
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), f"golden parameter recipe drifted for {n}: {got} vs {want}"
    return g, P


# ---------------------------------------------------------------------------------------------------------------------
# training-batch-scale fixture (tests/golden/make_golden_big.py): 1024 rays x 20 000 points, reference-generated
GOLDEN_BIG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_dtu20k_r1024.pt")
BIG_N_POINTS, BIG_N_RAYS = 20000, 1024


def image_gt(uv, seed=3):
    """An image-like colour target: smooth in the pixel position (what a photograph is at ray-batch scale)."""
    g = torch.Generator().manual_seed(seed)
    u = uv[0] / torch.tensor([512.0, 384.0])
    fr = torch.rand(3, 2, generator=g) * 4 + 1
    ph = torch.rand(3, generator=g) * 6.28
    return (0.5 + 0.45 * torch.sin((u[:, None, :] * fr[None]).sum(-1) * 3.0 + ph[None]))[None]


def big_inputs():
    """The deterministic inputs of the big fixture: scene, camera, pixels, mask target, RNG draws, colour targets."""
    from spurfies_b200 import scenes
    scene = scenes.dtu_like(BIG_N_POINTS, seed=24, radii=(0.3, 0.45))
    cam = scenes.camera(0, scene["cam_radius"])
    uv = (scenes.pixel_batch(BIG_N_RAYS, seed=7) - torch.tensor([256.0, 192.0])) * 0.45 + torch.tensor([256.0, 192.0])
    gt = scenes.synthetic_gt(BIG_N_RAYS, 7)
    state = torch.random.get_rng_state()
    torch.manual_seed(1234)   # the draws the reference makes from the global CPU generator, in its order
    rng = {"t_rand": torch.rand(BIG_N_RAYS, 128), "u": torch.rand(BIG_N_RAYS, 64), "sampling_idx": torch.randperm(128)[:32]}
    torch.random.set_rng_state(state)
    return scene, cam, uv, gt, rng, {"random": gt["rgb"], "image": image_gt(uv)}


def load_golden_big():
    """(golden, scene, cam, uv, gt, rng, targets, P): inputs regenerated from the recipe and verified by checksum."""
    g = torch.load(GOLDEN_BIG, weights_only=False)
    scene, cam, uv, gt, rng, targets = big_inputs()
    P = H.init_params(scene["pts"], scene["colors"], seed=g["recipe"]["param_seed"])
    P.neural_feats_geometry *= g["recipe"]["geometry_latent_scale"]
    P.neural_feats_color[:, 3:] *= g["recipe"]["color_latent_scale_from3"]
    got = {"pts": scene["pts"], "uv": uv, "t_rand": rng["t_rand"], "u": rng["u"], "sampling_idx": rng["sampling_idx"].float(),
           "gt_random": targets["random"], "gt_image": targets["image"], "neural_feats_color": P.neural_feats_color,
           "F_color.0": P.F_color[0][0], "R.0": P.R[0][0], "F_geometry.0": P.F_geometry[0][0]}
    for n, t in got.items():
        v, want = float(t.double().abs().sum()), g["checksum"][n]
        assert abs(v - want) <= 1e-9 * max(1.0, abs(want)), f"big golden recipe drifted for {n}: {v} vs {want}"
    return g, scene, cam, uv, gt, rng, targets, P


def big_grad_errors(got: dict, gold_grads: dict) -> dict:
    """max|a-b| / max|ref| per trainable tensor against a `gold[target]["grads"]` entry.  Latent tables are compared on
    the fixture's row subset, relative to the WHOLE reference table's max (stored), plus the relative error of the
    column sums and of the Frobenius norm over the subset."""
    out = {}
    for n, r in gold_grads.items():
        a = got[n].detach().double().cpu()
        if isinstance(r, dict):
            sub = a[::r["stride"]]
            ref = r["values"].double()
            out[n] = float((sub - ref).abs().max() / r["absmax"])
            out[n + " (rows, fro)"] = float((sub - ref).norm() / ref.norm())
            out[n + " (colsum)"] = float((a.sum(0) - r["colsum"].double()).abs().max() / r["colsum"].double().abs().max())
        else:
            out[n] = float((a - r.double()).abs().max() / r.double().abs().max())
    return out


def trainable(P):
    P.neural_feats_color.requires_grad_()
    P.neural_feats_geometry.requires_grad_()
    P.F_color = [(W.requires_grad_(), b.requires_grad_()) for W, b in P.F_color]
    P.R = [(W.requires_grad_(), b.requires_grad_()) for W, b in P.R]
    P.beta.requires_grad_()
    return P


def load_into_model(model, P):
    """Copy oracle parameters into the product PointVolSDF (same names as the reference state_dict)."""
    with torch.no_grad():
        model.neural_feats_color.copy_(P.neural_feats_color)
        model.neural_feats_geometry.copy_(P.neural_feats_geometry)
        for seq, layers in ((model.F_color, P.F_color), (model.F_geometry, P.F_geometry), (model.R, P.R)):
            lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
            for m, (W, b) in zip(lin, layers):
                m.weight.copy_(W)
                m.bias.copy_(b)
        model.T[0].weight.copy_(P.T[0])
        model.T[0].bias.copy_(P.T[1])
        model.density.beta.fill_(float(P.beta))
    for prm in list(model.F_geometry.parameters()) + list(model.T.parameters()):
        prm.requires_grad_(False)  # train.py:151-154
    return model


def rel_err(a, b):
    """max |a-b| / max |b|  -- the tolerance metric used throughout (1e-4 fp32 mode, 2e-2 bf16 mode)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
