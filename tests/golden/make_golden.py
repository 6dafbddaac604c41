"""Generate golden vectors by running the REFERENCE's own Python modules on CPU.

Runs only in the authoring container (needs /root/reference); the GPU box uses the committed
``tests/golden/*.pt``.  Nothing from the reference is copied: its modules are imported from
/root/reference with

  * ``torch_knnquery`` replaced by ``oracle.knn.RefVoxelGrid`` (the reference extension is CUDA-only,
    knnquery.py:5-8; its kernels are pinned separately, see oracle/knn.py header),
  * ``Tensor.cuda`` / ``Module.cuda`` / ``device="cuda"`` factory kwargs redirected to CPU
    (the model hard-codes .cuda(): density.py:19, ray_sampler.py:36-55, pointneus_disent.py:37-40),
  * I/O-only imports (imageio, skimage, plyfile, torch_scatter, GPUtil, omegaconf) stubbed,
  * ``_init_neural_info`` (reads ./data/*.ply, pointneus_disent.py:131-205) replaced by synthetic points.

Usage:  python tests/golden/make_golden.py
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("SPF_REFERENCE_ROOT", "/root/reference")


def install_shims():
    from oracle.knn import RefVoxelGrid
    m = types.ModuleType("torch_knnquery")
    m.VoxelGrid = RefVoxelGrid
    sys.modules["torch_knnquery"] = m
    for name in ("imageio", "skimage", "plyfile", "torch_scatter", "GPUtil", "omegaconf", "nvidia_smi"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["torch_scatter"].scatter_min = None
    sys.modules["torch_scatter"].scatter_mean = None
    sys.modules["omegaconf"].OmegaConf = object
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    for fn in ("zeros", "ones", "full", "empty", "tensor", "arange", "eye", "linspace", "rand", "randn",
               "zeros_like", "ones_like", "full_like"):
        orig = getattr(torch, fn)

        def wrap(*a, __orig=orig, **k):
            if "device" in k and str(k["device"]).startswith("cuda"):
                k["device"] = "cpu"
            return __orig(*a, **k)
        setattr(torch, fn, wrap)
    sys.path.insert(0, REF)


class Conf(dict):
    """Minimal pyhocon.ConfigTree duck-type (get_int/get_float/get_bool/get_list/get_config + attributes)."""
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return v
    def __setattr__(self, k, v):
        self[k] = v
    def _get(self, k, default=None):
        return self[k] if k in self else default
    get_int = get_float = get_bool = get_list = get_string = _get
    def get_config(self, k, default=None):
        v = self._get(k, default)
        return Conf(v) if isinstance(v, dict) and not isinstance(v, Conf) else v


def make_conf(near=0.5):
    return Conf(feature_vector_size=64, scene_bounding_sphere=3.0, initialize_colors=True, k=8, r=2, rbf=45,
                vox_res=300, max_shading_pts=80, white_bkgd=False,
                density=Conf(params_init=Conf(beta=0.1), beta_min=0.0001),
                ray_sampler=Conf(far=4.5, near=near, N_samples=64, N_samples_eval=128, N_samples_extra=32, eps=0.1,
                                 beta_iters=10, max_total_iters=5))


def build_reference_model(scene, seed=0):
    from spurfies.model import pointneus_disent as pd
    pts, colors = scene["pts"], scene["colors"]

    def fake_init(self):
        n = len(pts)
        self.register_buffer("neural_pts", pts.clone().float())
        self.register_parameter("neural_feats_color", torch.nn.Parameter(torch.empty(n, 64)))
        self.register_parameter("neural_feats_geometry", torch.nn.Parameter(torch.empty(n, 32)))
    pd.PointVolSDF._init_neural_info = fake_init
    torch.manual_seed(seed)
    model = pd.PointVolSDF(make_conf(), scan_id="24", dataset="dtu")
    return model


def params_from_oracle(model, P):
    """Load the oracle's seeded init into the reference model (same names/shapes as its state_dict)."""
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


def main():
    install_shims()
    from spurfies_b200 import scenes
    from oracle import hotpath as H
    from spurfies.model.loss import VolSDFLoss  # noqa: F401  (needs helpers.help -> loguru)

    scene = scenes.dtu_like(4000, seed=24, radii=(0.3, 0.45))
    # make the frozen prior / colour MLP outputs non-degenerate: scale latents up a little
    P = H.init_params(scene["pts"], scene["colors"], seed=0)
    P.neural_feats_geometry *= 8.0
    P.neural_feats_color[:, 3:] *= 500.0
    model = build_reference_model(scene)
    params_from_oracle(model, P)
    R = 48
    cam = scenes.camera(0, scene["cam_radius"])
    uv = scenes.pixel_batch(R, seed=7)
    # centre the batch on the object so most rays hit it
    uv = (uv - torch.tensor([256.0, 192.0])) * 0.45 + torch.tensor([256.0, 192.0])
    gt = scenes.synthetic_gt(R, 7)
    gold = {"scene": {"pts": scene["pts"], "colors": scene["colors"], "ranges": scene["ranges"]},
            # parameters are NOT stored (2.5 MB of MLP weights): tests rebuild them with
            # oracle.hotpath.init_params(seed=0) + the two scalings below and verify these checksums.
            "params_recipe": {"seed": 0, "geometry_latent_scale": 8.0, "color_latent_scale_from3": 500.0},
            "params_checksum": {n: float(t.double().abs().sum()) for n, t in
                                [("neural_feats_color", P.neural_feats_color), ("neural_feats_geometry", P.neural_feats_geometry),
                                 ("T.w", P.T[0]), ("beta", P.beta)]
                                + [(f"F_color.{i}", W) for i, (W, b) in enumerate(P.F_color)]
                                + [(f"F_geometry.{i}", W) for i, (W, b) in enumerate(P.F_geometry)]
                                + [(f"R.{i}", W) for i, (W, b) in enumerate(P.R)]},
            "uv": uv, "pose": cam["pose"], "intrinsics": cam["intrinsics"], "gt": gt}

    # ---- (1) point SDF queries: sdf_importance / get_sdf_eval (pointneus_disent.py:249-298, 348-421)
    g = torch.Generator().manual_seed(5)
    q = scene["pts"][torch.randperm(4000, generator=g)[:600]] + 0.02 * torch.randn(600, 3, generator=g)
    q = torch.cat([q, torch.rand(200, 3, generator=g) * 2 - 1], 0)
    model.eval()
    with torch.no_grad():
        gold["point_queries"] = q
        gold["sdf_importance"] = model.sdf_importance(q.clone())
        gold["get_sdf_eval"] = model.get_sdf_eval(q.clone())

    # ---- (2) sampler, training schedule (fast=1) with recorded RNG draws (ray_sampler.py:55, 514, 550)
    from spurfies.utils import rend_util
    model.train()
    ray_dirs, cam_loc = rend_util.get_camera_params(uv, cam["pose"], cam["intrinsics"])
    ray_dirs = ray_dirs.reshape(-1, 3)
    cam_loc_r = cam_loc.unsqueeze(1).repeat(1, R, 1).reshape(-1, 3)
    torch.manual_seed(1234)
    z_train, _ = model.ray_sampler.get_z_vals(ray_dirs, cam_loc_r, model, 1, 1)
    torch.manual_seed(1234)
    rng = {"t_rand": torch.rand(R, 128), "u": torch.rand(R, 64), "sampling_idx": torch.randperm(128)[:32]}
    gold["rng"] = rng
    gold["ray_dirs"], gold["cam_loc"] = ray_dirs, cam_loc
    gold["z_train"] = z_train
    # ---- (3) sampler, eval schedule (fast=-1, <=5 iterations, deterministic)
    model.eval()
    z_eval, _ = model.ray_sampler.get_z_vals(ray_dirs, cam_loc_r, model, -1, 1)
    gold["z_eval"] = z_eval

    # ---- (4) full training forward + loss + backward (train.py:330-397 without the optimiser)
    model.train()
    for prm in list(model.F_geometry.parameters()) + list(model.T.parameters()):
        prm.requires_grad_(False)  # train.py:151-154
    torch.manual_seed(1234)
    out = model({"intrinsics": cam["intrinsics"], "uv": uv, "pose": cam["pose"], "iter_step": 1, "local_data": None},
                fast=1)
    loss_fn = VolSDFLoss("torch.nn.L1Loss", local_weight=0.5, pseudo_weight=0.5, eikonal_weight=0.001,
                         rgb_weight=1.0, tv_weight=0.01)
    lo = loss_fn(out, gt)
    model.zero_grad()
    lo["loss"].backward()
    gold["train_out"] = {k: v.detach().clone() for k, v in out.items() if torch.is_tensor(v)}
    gold["train_loss"] = {k: v.detach().clone() for k, v in lo.items()}
    gr = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    gold["train_grads"] = gr
    # ---- (5) eval forward (fast=-1): normals, no jitter
    model.eval()
    out_e = model({"intrinsics": cam["intrinsics"], "uv": uv, "pose": cam["pose"], "iter_step": 1, "local_data": None},
                  fast=-1)
    gold["eval_out"] = {k: v.detach().clone() for k, v in out_e.items() if torch.is_tensor(v)}
    path = os.path.join(ROOT, "tests", "golden", "hotpath_dtu4k.pt")
    torch.save(gold, path)
    sz = os.path.getsize(path) / 1e6
    print(f"wrote {path} ({sz:.2f} MB)")
    print("train loss:", {k: float(v) for k, v in gold["train_loss"].items()})
    print("hit rays:", int((out["weights"].sum(-1) > 0).sum()), "/", R, " valid samples:", out["grad_theta"].shape[0])
    print("grad norms:", {k: float(v.norm()) for k, v in gr.items()})


if __name__ == "__main__":
    main()
