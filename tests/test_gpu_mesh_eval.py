"""GPU: the callers either side of the hot path (BASELINE configs 4 and 5) -- SDF volume for marching
cubes and the chunked full-image eval render -- against the oracle restatements.  fp32 mode, 1e-4 relative."""
import numpy as np
import pytest
import torch

from oracle import hotpath as H
from oracle import mesh as OM
from tests.helpers import load_golden, load_into_model, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def setup():
    from spurfies_b200.model import PointVolSDF, default_conf
    g, P = load_golden()
    model = PointVolSDF(default_conf(), "24", "dtu", neural_points=g["scene"]["pts"], neural_colors=g["scene"]["colors"])
    load_into_model(model, P)
    return g, P, model


GRID_PARAMS = np.array([[-0.45, -0.5, -0.6], [0.7, 0.72, 0.68]])   # bbs.npz-style [min; max] (eval_spurfies.py:142-149)


def test_sdf_volume_matches_reference_order_and_values(setup):
    from spurfies_b200 import mesh
    g, P, model = setup
    want, og = OM.surface_volume(P, P.make_grid(), GRID_PARAMS, resolution=24, splitn=7000)
    got = mesh.get_surface_by_grid(GRID_PARAMS, model, resolution=24, chunk=5000)   # several chunks, ragged tail
    for a, b in zip(got["xyz"], og["xyz"]):
        assert np.array_equal(a, b)
    vol = got["volume"].cpu().numpy()
    assert vol.shape == want.shape
    assert np.array_equal(vol == 1000.0, want == 1000.0)
    hit = want != 1000.0
    assert hit.sum() > 100 and (~hit).sum() > 100
    assert rel_err(torch.from_numpy(vol[hit]), torch.from_numpy(want[hit])) < TOL
    assert got["has_surface"] == (not (want.min() > 0 or want.max() < 0))
    assert abs(got["spacing"][0] - (og["xyz"][0][2] - og["xyz"][0][1])) < 1e-12


def test_sdf_volume_shards_without_collective(setup):
    from spurfies_b200 import mesh
    g, P, model = setup
    grid = mesh.get_grid_uniform(20, (-0.8, 0.8))
    whole, (lo, hi) = mesh.sdf_volume(model, grid["xyz"], chunk=3000)
    assert (lo, hi) == (0, 8000)
    parts = [mesh.sdf_volume(model, grid["xyz"], chunk=3000, rank=r, world=3) for r in range(3)]
    assert [p[1] for p in parts] == [(0, 2667), (2667, 5334), (5334, 8000)]
    assert torch.equal(torch.cat([p[0] for p in parts]), whole)
    # same values as the model's own point query on the reference's materialised grid points
    pts = OM.get_grid_uniform(20, (-0.8, 0.8))["grid_points"].cuda()
    assert torch.equal(whole, model.get_sdf_eval(pts))


def test_empty_and_all_masked_grids(setup):
    from spurfies_b200 import mesh
    g, P, model = setup
    far = mesh.get_grid_uniform(6, (5.0, 6.0))        # entirely outside the neural points' grid
    v, _ = mesh.sdf_volume(model, far["xyz"])
    assert v.numel() == 216 and bool((v == 1000.0).all())
    v1, r = mesh.sdf_volume(model, far["xyz"], rank=3, world=300)   # ranks beyond the work get empty slabs
    assert v1.numel() == r[1] - r[0] <= 1


def test_render_image_chunks_and_shards(setup):
    """Chunked / sharded eval render == the model called on the same pixel chunks (general.py:24-60 plumbing), and the
    single-chunk render == oracle eval forward."""
    from spurfies_b200 import eval as E
    from spurfies_b200 import scenes
    g, P, model = setup
    cam = scenes.camera(1, 2.3)
    uv = scenes.pixel_batch(96, seed=5)
    inp = {"uv": uv.cuda(), "pose": cam["pose"].cuda(), "intrinsics": cam["intrinsics"].cuda(), "local_data": None}
    one, (lo, hi) = E.render_image(model, inp, 96, n_pixels=96)
    assert (lo, hi) == (0, 96) and one["rgb_values"].shape == (96, 3) and one["weights"].shape == (96, 80)
    model.eval()
    # (not under no_grad: the oracle takes d sdf / d x with autograd.grad, pointneus_disent.py:315-323)
    ro = H.render_forward(P, P.make_grid(), uv, cam["pose"], cam["intrinsics"], H.SamplerCfg(), False, -1, None)
    for k in ("rgb_values", "weights", "depth_values", "normal_map"):   # tolerance of the eval sampler chain, as in
        e = rel_err(one[k].reshape(-1), ro[k].reshape(-1))              # test_gpu_hotpath.test_eval_forward_matches_reference
        assert e < 2e-2, (k, e)
    # rays that miss the cloud: their up-sampler rows are NaN in the reference as well, torch.sort puts NaN last
    # (ray_sampler.py:533, 559) -- same pattern here, every element written (this once depended on stale memory)
    z = model._last["z_vals"].cpu()
    assert int(torch.isnan(ro["z_vals"]).sum()) > 0
    assert torch.equal(torch.isnan(z), torch.isnan(ro["z_vals"]))
    assert rel_err(torch.nan_to_num(z), torch.nan_to_num(ro["z_vals"])) < 2e-2
    # shards: rank r of 2 renders its own contiguous half in 20-pixel chunks
    for r in range(2):
        part, (a, b) = E.render_image(model, inp, 96, n_pixels=20, rank=r, world=2)
        assert (a, b) == (48 * r, 48 * r + 48)
        chunks = []
        with torch.no_grad():
            for s in E.split_input(inp, 96, 20, a, b):
                chunks.append(model(s, aux_losses=False)["rgb_values"])
        assert torch.equal(part["rgb_values"], torch.cat(chunks))
    assert not model.training


def test_block_cyclic_volume_and_interleaved_render_equal_the_whole(setup):
    """Multi-GPU load balancing without a collective: the ranks' block-cyclic shares of the SDF grid scatter back to the
    volume of one rank bit for bit, and the interleaved pixel sets render the pixels the model renders for them."""
    from spurfies_b200 import eval as E
    from spurfies_b200 import mesh, scenes
    g, P, model = setup
    grid = mesh.get_grid_uniform(22, (-0.8, 0.8))
    whole, _ = mesh.sdf_volume(model, grid["xyz"], chunk=3000)
    back = torch.full_like(whole, float("nan"))
    for r in range(3):
        part, (lo, hi) = mesh.sdf_volume(model, grid["xyz"], chunk=1000, rank=r, world=3, cyclic_block=128)
        assert lo == 0 and hi == mesh.cyclic_local_count(22 ** 3, r, 3, 128) == part.numel()
        back[mesh.cyclic_global_index(torch.arange(hi), r, 3, 128).cuda()] = part
    assert torch.equal(back, whole)
    cam = scenes.camera(1, 2.3)
    uv = scenes.pixel_batch(96, seed=5)
    inp = {"uv": uv.cuda(), "pose": cam["pose"].cuda(), "intrinsics": cam["intrinsics"].cuda(), "local_data": None}
    covered = torch.zeros(96, dtype=torch.bool)
    for r in range(2):
        out, idx = E.render_image_interleaved(model, inp, 96, n_pixels=20, rank=r, world=2, block=8)
        assert torch.equal(idx, E.interleaved_pixels(96, r, 2, 8)) and out["rgb_values"].shape == (idx.numel(), 3)
        ref, _ = E.render_image(model, dict(inp, uv=uv[:, idx].cuda()), int(idx.numel()), n_pixels=20)
        assert torch.equal(out["rgb_values"], ref["rgb_values"]) and torch.equal(out["weights"], ref["weights"])
        covered[idx] = True
    assert bool(covered.all())


def test_pca_aligned_second_pass_matches_reference_restatement(setup):
    """get_surface_by_grid(higher_res=True, recon_pc=...) (plots.py:222-261): same principal axes and aligned grid as the
    oracle restatement, and the SDF at the ROTATED grid points (generated on the fly by the kernel) equals the oracle's on
    the materialised, rotated points -- except where a rotated point sits within fp32 rounding of an occupancy boundary."""
    from spurfies_b200 import mesh
    g, P, model = setup
    gen = torch.Generator().manual_seed(4)
    # stand-in for the 10 000 samples of the low-resolution mesh: points of the cloud with small noise, stretched so that
    # the principal axes are well separated
    pick = g["scene"]["pts"][torch.randperm(g["scene"]["pts"].shape[0], generator=gen)[:3000]]
    recon = pick * torch.tensor([1.0, 0.8, 0.6]) + 0.002 * torch.randn(3000, 3, generator=gen)
    want, og, vecs, s_mean, pts = OM.aligned_surface_volume(P, P.make_grid(), recon, resolution=24, splitn=7000)
    got = mesh.get_surface_by_grid(None, model, resolution=24, higher_res=True, chunk=5000, recon_pc=recon.cuda())
    assert float((got["vecs"].cpu() - vecs).abs().max()) < 1e-4 and float((got["mean"].cpu() - s_mean).abs().max()) < 1e-6
    for a, b in zip(got["xyz"], og["xyz"]):
        assert a.shape == b.shape and float(np.abs(a - b).max()) < 1e-5
    assert float((got["first_grid_point"].cpu() - pts[0]).abs().max()) < 1e-5
    vol = got["volume"].cpu().numpy()
    assert vol.shape == want.shape
    same_mask = (vol == 1000.0) == (want == 1000.0)
    assert same_mask.mean() > 0.999, same_mask.mean()
    hit = (want != 1000.0) & (vol != 1000.0)
    assert hit.sum() > 100
    # the grid points differ by fp32 rounding of the rotation (<= 1e-6), the SDF by its gradient times that
    assert rel_err(torch.from_numpy(vol[hit]), torch.from_numpy(want[hit])) < 1e-3
    with pytest.raises(ValueError):
        mesh.get_surface_by_grid(GRID_PARAMS, model, resolution=8, higher_res=True)


def test_graph_replayed_eval_chunks_equal_eager_chunks(setup):
    """render_image(graph=True): full-size chunks replay one captured CUDA graph (the eval forward has no host sync); the
    image is bit-identical to the eager render, for two different cameras through the same graphs (one per chunk size:
    full chunks and the ragged tail)."""
    from spurfies_b200 import eval as E
    from spurfies_b200 import scenes
    g, P, model = setup
    uv = scenes.pixel_batch(1100, seed=9)
    for view in (0, 2):
        cam = scenes.camera(view, 2.3)
        inp = {"uv": uv.cuda(), "pose": cam["pose"].cuda(), "intrinsics": cam["intrinsics"].cuda(), "local_data": None}
        eager, _ = E.render_image(model, inp, 1100, n_pixels=256)
        graphed, _ = E.render_image(model, inp, 1100, n_pixels=256, graph=True)      # 4 replays + the 76-pixel tail's own graph
        for k in E.RENDER_KEYS:
            assert torch.equal(torch.nan_to_num(graphed[k]), torch.nan_to_num(eager[k])), k
    graphs = [v for v in model._eval_graphs.values()]
    assert len(graphs) == 2 and all(v is not None for v in graphs)   # captured once per chunk size, reused by the second camera
