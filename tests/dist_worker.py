"""Worker of tests/test_gpu_dist.py: run under ``torchrun --nproc-per-node N`` on a multi-GPU box (NCCL over NVLink).

SURVEY 7 T4 / 8(e): a ray-sharded data-parallel training step must equal the one-big-batch step.  Every rank builds
two identical models; one takes the sharded step through TrainStep(world_size=N) (global-count loss normalisation,
one NCCL all-reduce of the flat gradient, identical Adam step on every rank), the other the whole batch through
TrainStep(world_size=1).  Rank 0 prints one JSON line with the loss terms, the gradient and post-Adam parameter
differences."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    from spurfies_b200 import scenes
    from spurfies_b200.dist import shard_range
    from spurfies_b200.model import PointVolSDF, default_conf
    from spurfies_b200.train import TrainStep
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = scenes.dtu_like(8000, seed=1, radii=(0.35, 0.5))
    R = 512
    cam = scenes.camera(0, sc["cam_radius"])
    uv = (scenes.pixel_batch(R, 3) - torch.tensor([256.0, 192.0])) * 0.8 + torch.tensor([256.0, 192.0])  # some rays miss
    # rays ordered by distance from the image centre: the first shard hits the object with (almost) every ray, the last
    # one misses with most -> the count-normalised loss terms see very different denominators on the ranks
    uv = uv[:, torch.argsort((uv[0] - torch.tensor([256.0, 192.0])).norm(dim=-1))]
    gt, rng = scenes.synthetic_gt(R, 3), scenes.rng_inputs(R, 3)
    ld = scenes.local_data(0, sc["cam_radius"], feat_res=(128, 96))

    def make(w, **kw):
        torch.manual_seed(0)
        m = PointVolSDF(default_conf(), "24", "dtu", neural_points=sc["pts"], neural_colors=sc["colors"], precision=precision)
        with torch.no_grad():
            m.neural_feats_geometry.mul_(8.0)
            m.neural_feats_color[:, 3:].mul_(500.0)
        return TrainStep(m, world_size=w, lr_schedule=False, **kw)

    cu = lambda d: {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()}

    def inputs(lo, hi):
        b = {"uv": uv[:, lo:hi].cuda(), "pose": cam["pose"].cuda(), "intrinsics": cam["intrinsics"].cuda(), "local_data": cu(ld)}
        g = {"rgb": gt["rgb"][:, lo:hi].cuda(), "mask": gt["mask"][:, lo:hi].cuda()}
        r = {"t_rand": rng["t_rand"][lo:hi].cuda(), "u": rng["u"][lo:hi].cuda(), "sampling_idx": rng["sampling_idx"].cuda()}
        return b, g, r

    big, dp = make(1), make(world)
    p0 = big.opt.flat_p.clone()
    assert torch.equal(p0, dp.opt.flat_p)
    # gradients: run forward + backward (+ all-reduce) by hand so the flat gradient can be read before Adam clears it
    def grads(step, b, g, r):
        step.model.train()
        out = step.model(b, fast=1, rng=r, dense_outputs=True)
        lo_ = step.loss(out, g)
        if not step._reducer.attached():
            step._reducer.attach()
            step._reducer.flat().zero_()
        for p in step.params:
            p._spf_direct_grad = True
        lo_["loss"].backward()
        step._allreduce_grads()
        return lo_, step._reducer.flat().clone() / step.world_size

    lo, hi = shard_range(R, rank, world)
    l_big, g_big = grads(big, *inputs(0, R))
    l_dp, g_dp = grads(dp, *inputs(lo, hi))
    terms = ["loss", "rgb_loss", "mask_loss", "pseudo_loss", "eikonal_loss", "local_loss", "tv_loss"]
    t_dp = torch.stack([l_dp[k].detach().float().reshape(()) for k in terms])
    dist.all_reduce(t_dp)
    t_dp /= world
    t_big = torch.stack([l_big[k].detach().float().reshape(()) for k in terms])
    # hit counts differ between the shards (otherwise the global normalisation would not be exercised)
    nv = torch.zeros(world, device="cuda")
    nv[rank] = float((dp.model._last["ray_mask"]).sum())
    dist.all_reduce(nv)
    # the optimiser step on both
    big.opt.step(grad_scale=1.0, zero_grad=True)
    dp.opt.step(grad_scale=1.0 / world, zero_grad=True)
    d_big, d_dp = big.opt.flat_p - p0, dp.opt.flat_p - p0
    # every rank must hold identical parameters after the step
    chk = torch.stack([dp.opt.flat_p.double().sum(), dp.opt.flat_p.double().abs().sum()])
    lo_chk, hi_chk = chk.clone(), chk.clone()
    dist.all_reduce(lo_chk, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_chk, op=dist.ReduceOp.MAX)
    # the whole step through TrainStep.__call__: early all-reduce of the colour-latent gradient inside the backward
    # (dp_overlap, opt-in) against the single all-reduce after it, eagerly and as a replayed CUDA graph
    variants = {}
    for name, kw, graph in (("overlap", {"dp_overlap": True}, False), ("serial", {"dp_overlap": False}, False),
                            ("overlap_graph", {"dp_overlap": True}, True)):
        st = make(world, **kw)
        assert (st._early_n > 0) == kw["dp_overlap"]
        if graph:
            captured = st.capture(*inputs(lo, hi))
            assert captured, st.graph_error
        st(*inputs(lo, hi))
        variants[name] = (st.opt.flat_p - p0).clone()
    torch.cuda.synchronize()
    if rank == 0:
        gmax = float(g_big.abs().max())
        print(json.dumps({
            "world": world, "precision": precision, "hit_rays_per_rank": [int(v) for v in nv],
            "terms": {k: [float(a), float(b)] for k, a, b in zip(terms, t_dp, t_big)},
            "grad_rel_err": float((g_dp - g_big).abs().max() / gmax), "grad_absmax": gmax,
            "adam_delta_absmax": float(d_big.abs().max()),
            "adam_delta_mean_abs_diff": float((d_dp - d_big).abs().mean()), "adam_delta_max_abs_diff": float((d_dp - d_big).abs().max()),
            "adam_delta_mean_abs": float(d_big.abs().mean()),
            "overlap_vs_serial_mean_abs_diff": float((variants["overlap"] - variants["serial"]).abs().mean()),
            "graph_vs_eager_mean_abs_diff": float((variants["overlap_graph"] - variants["overlap"]).abs().mean()),
            "overlap_vs_manual_mean_abs_diff": float((variants["overlap"] - d_dp).abs().mean()),
            "ranks_identical": bool(torch.equal(lo_chk, hi_chk))}))
        sys.stdout.flush()
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
