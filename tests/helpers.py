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
