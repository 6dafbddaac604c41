"""Development: fraction of real (sample, neighbour) pairs among the 8 rows per valid slot of the bench step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from spurfies_b200.train import TrainStep
dev = torch.device("cuda", 0)
for wl in ("train", "garden"):
    w = bench.WORKLOADS[wl]
    sc, model = bench.build_scene(dev, scene=w["scene"], n_points=w["n_points"])
    step = TrainStep(model)
    hb = bench.host_batches(2, 0, n_rays=4096, cam_radius=sc["cam_radius"])
    for h in hb:
        b, g, r = bench.split(bench.to_device(h, dev))
        step(b, g, r)
    torch.cuda.synchronize()
    s = model._last["slots"]
    V = s.V
    pv = s.pidx[s.list[:V].long()]
    real = int((pv >= 0).sum())
    hist = torch.bincount((pv >= 0).sum(1), minlength=9).tolist()
    print(wl, "valid slots", V, "rows", V * 8, "real pairs", real, "fraction", real / (V * 8), "hist of neighbours/slot", hist)
