"""Debug aid: where does the eval render of tests/test_gpu_mesh_eval.py::test_render_image_chunks_and_shards go NaN?
Replays the module's tests in order (the failure depends on what earlier tests left in the caching allocator)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import test_gpu_mesh_eval as T
from tests.helpers import load_golden, load_into_model
from spurfies_b200 import scenes
from spurfies_b200.model import PointVolSDF, default_conf

g, P = load_golden()
model = load_into_model(PointVolSDF(default_conf(), "24", "dtu", neural_points=g["scene"]["pts"], neural_colors=g["scene"]["colors"]), P)
setup = (g, P, model)
T.test_sdf_volume_matches_reference_order_and_values(setup)
T.test_sdf_volume_shards_without_collective(setup)
T.test_empty_and_all_masked_grids(setup)
model.eval()
cam = scenes.camera(1, 2.3)
uv = scenes.pixel_batch(96, seed=5)
inp = {"uv": uv.cuda(), "pose": cam["pose"].cuda(), "intrinsics": cam["intrinsics"].cuda(), "local_data": None}
for rep in range(3):
    with torch.no_grad():
        out = model(inp, fast=-1, aux_losses=False)
    L = model._last
    print("rep", rep)
    for k in ("rgb_values", "weights", "depth_values", "normal_map", "xyz", "depth_vals"):
        v = out[k]
        bad = torch.isnan(v).reshape(v.shape[0], -1).any(1)
        print(f"  out[{k}]: nan rows {bad.nonzero().flatten().tolist()[:12]}")
    for k in ("z_vals", "sdf", "delta", "t", "rgb_s", "dist", "acc", "loc"):
        v = L[k]
        print(f"  last[{k}] shape {tuple(v.shape)} nan {int(torch.isnan(v).sum())} inf {int(torch.isinf(v).sum())}")
    pidx = L["slots"].pidx.view(96, 80, -1)
    valid = pidx[..., 0] >= 0
    sdf = L["sdf"].view(96, 80); rgb_s = L["rgb_s"].view(96, 80, 3)
    print("  nan sdf on valid slots:", int((torch.isnan(sdf) & valid).sum()), " nan rgb_s on valid:", int((torch.isnan(rgb_s).any(-1) & valid).sum()),
          " valid slots:", int(valid.sum()), "rays:", int(L["ray_mask"].sum()), "count", int(L["slots"].count))
    for nm, bad in (("rgb_s", (torch.isnan(rgb_s).any(-1) & valid).nonzero()), ("sdf", (torch.isnan(sdf) & valid).nonzero())):
        print("  first bad", nm, "(ray, slot):", bad[:8].tolist())
        if len(bad):
            r, s = bad[0].tolist()
            print("  pidx", pidx[r, s].tolist(), "loc", L["loc"].view(96, 80, 3)[r, s].tolist(), "xyz", out["xyz"][r, s].tolist(), "sdf", float(sdf[r, s]),
                  "t", float(L["t"][r, s]), "delta", float(L["delta"][r, s]), "ray_dir", L["ray_dirs"][r].tolist(), "nvalid slots in ray", int(valid[r].sum()))
