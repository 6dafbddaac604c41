"""Development: a small end-to-end workload for `compute-sanitizer --tool memcheck` (smoke() + two bf16 training steps
+ one eval chunk + one mesh slab on a 20 k-point scene)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
import bench
from spurfies_b200 import mesh
from spurfies_b200.train import TrainStep

g.smoke()
dev = torch.device("cuda", 0)
sc, model = bench.build_scene(dev, precision="bf16", n_points=20000)
step = TrainStep(model)
hb = bench.host_batches(2, 0, n_rays=512, cam_radius=sc["cam_radius"])
for h in hb:
    b, gt, r = bench.split(bench.to_device(h, dev))
    print("loss", float(step(b, gt, r)["loss"]))
model.eval()
with torch.no_grad():
    b, gt, r = bench.split(bench.to_device(hb[0], dev))
    out = model(b, fast=-1)
    print("eval rgb", float(out["rgb_values"].mean()))
    grid = mesh.get_grid_uniform(48, (-1.0, 1.0))
    vol = mesh.sdf_volume(model, grid["xyz"], chunk=1 << 16)
    vol = vol[0] if isinstance(vol, tuple) else vol
    print("mesh", tuple(vol.shape), float((vol < 999).float().mean()))
torch.cuda.synchronize()
print("done")
