"""Development: where does the full-size eval render differ from the oracle?  (run on the GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import hotpath as H
from spurfies_b200 import scenes, eval as E
from tests.test_gpu_baseline_sizes import CONFIGS, _scene_params, _model
from tests.helpers import rel_err

cfg = CONFIGS["configs1_dtu_100k_4096rays"]
sc, P = _scene_params(cfg)
model = _model(cfg, sc, P, "fp32")
cam = scenes.camera(1, sc["cam_radius"])
uv = scenes.full_image_uv()
n = uv.shape[1]
inp = {"uv": uv.cuda(), "pose": cam["pose"].cuda(), "intrinsics": cam["intrinsics"].cuda(), "local_data": None}
chunk = 6
part, (a, b) = E.render_image(model, inp, n, n_pixels=16384, rank=chunk, world=n // 16384)
z_chunk = model._last["z_vals"].cpu()
print("chunk iters", int(model.ray_sampler.last_iters_used.item()))
loc = (torch.arange(256) * 61 + 17) % 16384
sub = a + loc
inp_s = dict(inp, uv=uv[:, sub].cuda())
small, _ = E.render_image(model, inp_s, 256, n_pixels=256)
z_small = model._last["z_vals"].cpu()
print("small iters", int(model.ray_sampler.last_iters_used.item()))
print("product chunk vs product small: z", rel_err(torch.nan_to_num(z_chunk[loc]), torch.nan_to_num(z_small)),
      {k: rel_err(part[k][loc.cuda()], small[k]) for k in ("rgb_values", "weights")})
scfg = H.SamplerCfg()
ray_dirs, cam_loc = H.camera_rays(uv[:, sub], cam["pose"], cam["intrinsics"])
ray_dirs = ray_dirs.reshape(-1, 3)
cl = cam_loc.unsqueeze(1).repeat(1, ray_dirs.shape[0], 1).reshape(-1, 3)
for fi in (None, 1, 2, 3, 4, 5):
    tr = {}
    z = H.sample_z(P, P.make_grid(), ray_dirs, cl, scfg, False, -1, None, force_iters=fi, trace=tr)
    dz_c = (torch.nan_to_num(z) - torch.nan_to_num(z_chunk[loc])).abs().max(dim=1).values
    dz_s = (torch.nan_to_num(z) - torch.nan_to_num(z_small)).abs().max(dim=1).values
    print("oracle force", fi, "ran", len(tr["iters"]), "max|dz| vs chunk", float(dz_c.max()), "rays off > 1e-4:", int((dz_c > 1e-4).sum()),
          "| vs small", float(dz_s.max()), int((dz_s > 1e-4).sum()))
