"""CPU study (oracle only, no GPU): how far is a reduced-precision-operand evaluation of the training step from the
fp32 one, per trainable tensor, as a function of (a) the operand rounding (bf16 / fp16 / bf16 hi+lo), (b) the batch
size and (c) how coherent the colour target is.  Explains the tolerance of tests/test_gpu_hotpath.py::test_bf16_mode_*.

usage: python tools/bf16_grad_study.py [R=1024] [N=20000]
"""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hotpath as H  # noqa: E402
from spurfies_b200 import scenes  # noqa: E402
from tests.helpers import rel_err, trainable  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20000

ROUND = {"fn": None, "geo": False}


def r_bf16(t):
    return t.to(torch.bfloat16).float()


def r_fp16(t):
    return t.to(torch.float16).float()


def r_split(t):
    hi = t.to(torch.bfloat16).float()
    return hi + (t - hi).to(torch.bfloat16).float()


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return ROUND["fn"](t)

    @staticmethod
    def backward(ctx, g):
        return ROUND["bwd"](g) if ROUND.get("bwd") else g


_orig_mlp = H.mlp


def mlp_rounded(x, layers, act_last=False):
    if ROUND["fn"] is None or (not ROUND["geo"] and layers[0][0].shape[1] == 35):
        return _orig_mlp(x, layers, act_last)
    n = len(layers)
    for i, (W, b) in enumerate(layers):
        x = F.linear(_Round.apply(x), _Round.apply(W), b)
        if i < n - 1 or act_last:
            x = F.leaky_relu(x, H.LEAKY)
    return x


H.mlp = mlp_rounded


def smooth_gt(uv, seed):
    """An image-like target: smooth in the pixel position (what a photograph is at ray-batch scale)."""
    g = torch.Generator().manual_seed(seed)
    u = uv[0] / torch.tensor([512.0, 384.0])
    fr = torch.rand(3, 2, generator=g) * 4 + 1
    ph = torch.rand(3, generator=g) * 6.28
    rgb = 0.5 + 0.45 * torch.sin((u[:, None, :] * fr[None]).sum(-1) * 3.0 + ph[None])
    return rgb[None]


FIXED_Z = {"z": None}   # FIX_Z=1: every variant renders at the fp32 run's sample positions (isolates the sampler's
                        # discrete effects -- a sample crossing a voxel / neighbour-set boundary -- from the MLP arithmetic)


def run(P, sc, uv, cam, rng, gt_rgb, gt_mask):
    Pt = trainable(P)
    ro = H.render_forward(Pt, Pt.make_grid(), uv, cam["pose"], cam["intrinsics"], H.SamplerCfg(), True, 1, rng,
                          z_vals=FIXED_Z["z"])
    if os.environ.get("FIX_Z") == "1" and FIXED_Z["z"] is None:
        FIXED_Z["z"] = ro["z_vals"].detach().clone()
    lo = H.volsdf_loss(ro, gt_rgb, gt_mask)
    lo["loss"].backward()
    g = {"feat_c": Pt.neural_feats_color.grad, "feat_g": Pt.neural_feats_geometry.grad, "beta": Pt.beta.grad}
    for nm, layers in (("F_color", Pt.F_color), ("R", Pt.R)):
        for i, (W, b) in enumerate(layers):
            g[f"{nm}.{2 * i}.w"], g[f"{nm}.{2 * i}.b"] = W.grad, b.grad
    return {k: v.clone() for k, v in g.items()}, float(lo["loss"])


def main():
    sc = scenes.dtu_like(N, seed=24, radii=(0.3, 0.45))
    cam = scenes.camera(0, sc["cam_radius"])
    uv = (scenes.pixel_batch(R, seed=7) - torch.tensor([256.0, 192.0])) * 0.45 + torch.tensor([256.0, 192.0])
    rng = scenes.rng_inputs(R, step=1)
    gt = scenes.synthetic_gt(R, 7)
    targets = {"random": gt["rgb"], "smooth": smooth_gt(uv, 3)}

    def fresh():
        P = H.init_params(sc["pts"], sc["colors"], seed=0)
        P.neural_feats_geometry *= 8.0
        P.neural_feats_color[:, 3:] *= 500.0
        return P
    for tn, tg in targets.items():
        ROUND["fn"] = None
        t0 = time.time()
        ref, l0 = run(fresh(), sc, uv, cam, rng, tg, gt["mask"][0, :, 0])
        print(f"[{tn}] fp32 loss {l0:.6f} ({time.time() - t0:.1f}s)", flush=True)
        variants = (("bf16", r_bf16, False, None), ("bf16+bwd", r_bf16, False, r_bf16),
                    ("fp16", r_fp16, False, None), ("split", r_split, False, None),
                    ("split+bf16bwd", r_split, False, r_bf16), ("bf16 geo too", r_bf16, True, None),
                    ("fp16 geo too+bf16bwd", r_fp16, True, r_bf16), ("split geo too", r_split, True, None),
                    ("split geo too+bf16bwd", r_split, True, r_bf16))
        only = os.environ.get("VARIANTS")
        for name, fn, geo, bwd in variants:
            if only and name not in only.split(","):
                continue
            ROUND.update(fn=fn, geo=geo, bwd=bwd)
            got, l1 = run(fresh(), sc, uv, cam, rng, tg, gt["mask"][0, :, 0])
            e = {k: rel_err(got[k], ref[k]) for k in ref}
            worst = max(e.values())
            print(f"[{tn}] {name:14s} loss {l1:.6f} worst {worst:.2e} :: " + " ".join(f"{k}={v:.1e}" for k, v in e.items()),
                  flush=True)


if __name__ == "__main__":
    main()
