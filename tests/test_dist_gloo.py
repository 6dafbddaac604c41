"""CPU, world_size 2 over gloo: the host-side logic of the ray-sharded data-parallel step (spurfies_b200/dist.py).

The reference is single-process; the contract is "same result as one big batch": with the step's rays
split into equal contiguous shards, the all-reduced (averaged) gradient of the per-ray loss terms equals the
gradient of the full batch.  The per-ray arithmetic here is the oracle (CPU checker), the reduction is the product's
FlatGradReducer -- exactly what TrainStep runs between backward and the clip / Adam step on the GPU box.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from spurfies_b200.dist import FlatGradReducer, max_over_ranks, shard_range, shard_rays  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 4096, 196608, 134217728):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
    uv = torch.arange(10.0).reshape(1, 5, 2)
    parts = [shard_rays({"uv": uv, "pose": 1}, r, 2)["uv"] for r in range(2)]
    assert torch.equal(torch.cat(parts, 1), uv)


def _worker_reduce(rank, world, port, q):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(()))]
    params[0].grad = torch.full((5, 3), float(rank + 1))
    params[1].grad = None if rank == 0 else torch.arange(7.0)      # a missing grad counts as zero
    params[2].grad = torch.tensor(10.0 * rank)
    red = FlatGradReducer(params, world)
    red.reduce()
    ok = (torch.allclose(params[0].grad, torch.full((5, 3), 1.5)) and torch.allclose(params[1].grad, torch.arange(7.0) / 2)
          and torch.allclose(params[2].grad, torch.tensor(5.0)) and red.bytes_per_step == 4 * (15 + 7 + 1))
    first = red.flat().data_ptr()
    red.reduce()
    ok = ok and red.flat().data_ptr() == first                      # persistent buffer (CUDA-graph capturable)
    # attached mode: p.grad are views of the flat buffer, autograd accumulates into it, reduce() needs no packing
    w = [torch.nn.Parameter(torch.ones(4, 2) * (rank + 1)), torch.nn.Parameter(torch.ones(3))]
    red2 = FlatGradReducer(w, world)
    flat = red2.attach()
    (w[0].pow(2).sum() + (3.0 * w[1]).sum()).backward()
    ok = ok and red2.attached() and torch.allclose(flat[:8], torch.full((8,), 2.0 * (rank + 1)))
    red2.reduce()
    ok = ok and red2.attached() and torch.allclose(w[0].grad, torch.full((4, 2), 3.0)) and torch.allclose(w[1].grad, torch.full((3,), 3.0))
    red2.zero()
    ok = ok and float(w[0].grad.abs().sum()) == 0.0
    mx = max_over_ranks([float(rank), 3.0 - rank], "cpu")
    ok = ok and mx == [1.0, 3.0]
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_flat_grad_reducer_world2():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_reduce, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


def _worker_reduce_bf16(rank, world, port, q):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(17)
    base = [torch.randn(300, 8, generator=g), torch.randn(300, 4, generator=g), torch.randn(16, 16, generator=g),
            torch.randn(5, generator=g)]
    w = [torch.nn.Parameter(torch.zeros_like(b)) for b in base]
    n_lat = 300 * 8 + 300 * 4
    red = FlatGradReducer(w, world, align=4, bf16_prefix=n_lat)   # the two "latent tables" travel as bf16
    red.attach()
    for p, b in zip(w, base):
        p.grad.copy_(b * (rank + 1))
    try:
        red.reduce(average=False)
    except RuntimeError as e:          # a gloo build without bf16 reductions
        q.put((rank, "unsupported: " + str(e)[:80]))
        dist.destroy_process_group()
        return
    want = [b * 3.0 for b in base]     # ranks contribute 1x and 2x
    ok = True
    for i, (p, t) in enumerate(zip(w, want)):
        if i < 2:                      # bf16 round trip: 2^-8 relative per element, exact structure
            ok = ok and bool(((p.grad - t).abs() <= 2.0 ** -7 * t.abs() + 1e-6).all()) and not torch.equal(p.grad, t)
        else:                          # the fp32 tail is exact
            ok = ok and torch.equal(p.grad, t)
    ok = ok and red.bytes_per_step == 4 * red.numel - 2 * n_lat
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_flat_grad_reducer_bf16_prefix_world2():
    """Opt-in compression of the latent-table part of the gradient exchange: bf16 for the prefix, fp32 for the rest."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_reduce_bf16, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    if any(isinstance(r[1], str) for r in res):
        pytest.skip(str(res))
    assert res == [(0, True), (1, True)]


def _oracle_setup():
    from oracle import hotpath as H
    from spurfies_b200 import scenes
    sc = scenes.dtu_like(6000, seed=3, radii=(0.35, 0.5))
    P = H.init_params(sc["pts"], sc["colors"], seed=3)
    P.neural_feats_geometry *= 8.0
    P.neural_feats_color[:, 3:] *= 500.0
    for t in P.trainable():
        t.requires_grad_()
    cam = scenes.camera(0, sc["cam_radius"])
    R = 32
    uv = (scenes.pixel_batch(R, seed=5) - torch.tensor([256.0, 192.0])) * 0.5 + torch.tensor([256.0, 192.0])
    rng, gt = scenes.rng_inputs(R, step=5), scenes.synthetic_gt(R, 5)
    return H, P, cam, uv, rng, gt


def _per_ray_loss(H, P, grid, cam, uv, rng, gt_rgb, gt_mask, world=1):
    """rgb + mask (means over the rays: equal shards combine exactly) + the pseudo-point term, a mean over the HIT rays
    whose pseudo point has a neighbour -- a data-dependent count that differs between shards, normalised globally with
    the same helper PointVolSDF.forward uses (spurfies_b200/dist.py::global_count_scales)."""
    from spurfies_b200.dist import global_count_scales
    out = H.render_forward(P, grid, uv, cam["pose"], cam["intrinsics"], H.SamplerCfg(), True, 1, rng, with_tv=False)
    lo = H.volsdf_loss(out, gt_rgb, gt_mask)
    cnt = torch.tensor([float(out.get("pseudo_count", 0))])
    scale = global_count_scales(cnt, world)[0]
    return lo["rgb_loss"] + lo["mask_loss"] + 0.5 * lo["pseudo_loss"] * scale, int(cnt[0])


def _worker_sharded(rank, world, port, q):
    torch.set_num_threads(2)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    H, P, cam, uv, rng, gt = _oracle_setup()
    R = uv.shape[1]
    lo, hi = shard_range(R, rank, world)
    rng_r = {"t_rand": rng["t_rand"][lo:hi], "u": rng["u"][lo:hi], "sampling_idx": rng["sampling_idx"]}
    loss, cnt = _per_ray_loss(H, P, P.make_grid(), cam, shard_rays({"uv": uv}, rank, world)["uv"], rng_r,
                              gt["rgb"][:, lo:hi], gt["mask"][0, lo:hi, 0], world)
    loss.backward()
    params = P.trainable()
    FlatGradReducer(params, world).reduce()
    q.put((rank, ([p.grad.detach().numpy().copy() for p in params], cnt)))  # by value (no fd passing)
    dist.destroy_process_group()


def test_sharded_rays_equal_one_big_batch_world2():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(60)
    H, P, cam, uv, rng, gt = _oracle_setup()
    loss, cnt = _per_ray_loss(H, P, P.make_grid(), cam, uv, rng, gt["rgb"], gt["mask"][0, :, 0])
    loss.backward()
    want = [p.grad for p in P.trainable()]
    nonzero = 0
    # the shards see different hit counts (otherwise the test would not exercise the global normalisation)
    assert got[0][1] + got[1][1] == cnt and got[0][1] != got[1][1], (got[0][1], got[1][1], cnt)
    for g0, g1, w in zip(got[0][0], got[1][0], want):
        g0, g1 = torch.from_numpy(g0), torch.from_numpy(g1)
        assert torch.equal(g0, g1)                                    # every rank holds the identical reduced gradient
        scale = float(w.abs().max())
        nonzero += scale > 0
        assert float((g0 - w).abs().max()) <= 2e-5 * max(scale, 1e-12) + 1e-9, (float((g0 - w).abs().max()), scale)
    assert nonzero >= 10
