"""GPU: the reference's GPU path on the same device (oracle/gpu_ref.py) -- its own compiled CUDA kernels under the
restated torch graph -- and the Level-1 drop-in check: the SAME reference-side graph, glue included
(set_pointset before every query, utils.query's ragged compaction: spurfies/model/utils.py:90-113), run once on top of
the reference's `torch_knnquery.VoxelGrid` kernels and once on top of `spurfies_b200.knnquery.VoxelGrid`, must give the
same training step."""
import pytest
import torch

from oracle import hotpath as H
from tests.helpers import rel_err, trainable

pytestmark = pytest.mark.gpu


def _setup(n_points=20000, R=256):
    from oracle import gpu_ref as G
    from spurfies_b200 import scenes
    if G.load_reference_ext() is None:
        pytest.skip("oracle/_ref/knnquery_cuda*.so not built (needs /root/reference at build time)")
    sc = scenes.dtu_like(n_points, seed=11, radii=(0.4, 0.6))
    cam = scenes.camera(2, sc["cam_radius"])
    cam = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in cam.items()}
    uv = ((scenes.pixel_batch(R, seed=5) - torch.tensor([256.0, 192.0])) * 0.5 + torch.tensor([256.0, 192.0])).cuda()
    rng = {k: v.cuda() for k, v in scenes.rng_inputs(R, step=3).items()}
    gt = {k: v.cuda() for k, v in scenes.synthetic_gt(R, 5).items()}

    def params():
        P = H.init_params(sc["pts"], sc["colors"], seed=3)
        P.neural_feats_geometry *= 8.0
        P.neural_feats_color[:, 3:] *= 500.0
        return trainable(G.params_to(P, "cuda"))
    return G, sc, cam, uv, rng, gt, params


def test_reference_graph_runs_unchanged_on_the_product_voxelgrid():
    G, sc, cam, uv, rng, gt, params = _setup()
    res = {}
    for kind in ("reference", "product"):
        P = params()
        grid = G.make_grid(kind, P)
        out, lo = G.training_step(P, grid, uv, cam, rng, gt)
        torch.cuda.synchronize()
        res[kind] = (out, lo, [t.grad.clone() for t in P.trainable()])
    (o_r, l_r, g_r), (o_p, l_p, g_p) = res["reference"], res["product"]
    assert int(o_r["ray_mask"].sum()) > 100 and torch.equal(o_r["ray_mask"], o_p["ray_mask"])
    assert torch.equal(o_r["mask"], o_p["mask"])                          # same slots valid
    assert torch.equal(o_r["neighbor_idx"], o_p["neighbor_idx"])          # same neighbour sets (sorted by id in the adapter)
    errs = {k: rel_err(o_p[k], o_r[k]) for k in ("rgb_values", "weights", "depth_values", "xyz")}
    errs.update({"loss." + k: abs(float(l_p[k]) - float(l_r[k])) / max(1.0, abs(float(l_r[k]))) for k in l_r})
    errs["grads"] = max(rel_err(a, b) for a, b in zip(g_p, g_r) if float(b.abs().max()) > 0)
    print("reference graph on product VoxelGrid vs on reference kernels:", {k: f"{v:.1e}" for k, v in errs.items()})
    # identical inputs to an identical torch graph: only cuBLAS / atomics run-to-run noise remains.  The gradients are
    # max-norm errors of scatter-added sums whose order changes from run to run (torch index_add atomics; the reference
    # kernels emit each neighbour list in a nondeterministic order): observed 4e-5 ... 1.4e-4 over repeated runs
    grads = errs.pop("grads")
    assert max(errs.values()) < 1e-4 and grads < 5e-4, (errs, grads)


def test_reference_gpu_path_matches_the_product_kernels():
    """The reference-side GPU step (its kernels + torch graph) against the product's exact (fp32) mode at IDENTICAL
    sample depths (the product's z injected into the reference graph; the sampler is compared separately below):
    rendered outputs and loss within 1e-4, gradients within 1e-3."""
    from spurfies_b200.model import PointVolSDF, VolSDFLoss, default_conf
    from tests.helpers import load_into_model
    G, sc, cam, uv, rng, gt, params = _setup()
    Pc = H.init_params(sc["pts"], sc["colors"], seed=3)
    Pc.neural_feats_geometry *= 8.0
    Pc.neural_feats_color[:, 3:] *= 500.0
    model = load_into_model(PointVolSDF(default_conf(), "24", "dtu", neural_points=sc["pts"], neural_colors=sc["colors"],
                                        max_points_per_voxel=128, max_occ_voxels=32768), Pc)
    model.train()
    mo = model({"intrinsics": cam["intrinsics"], "uv": uv, "pose": cam["pose"], "local_data": None}, fast=1, rng=rng)
    ml = VolSDFLoss()(mo, gt)
    ml["loss"].backward()
    P = params()
    out, lo = G.training_step(P, G.make_grid("reference", P), uv, cam, rng, gt, z_vals=model._last["z_vals"].clone())
    e = {k: rel_err(mo[k], out[k]) for k in ("rgb_values", "weights", "depth_values")}
    e["loss"] = abs(float(ml["loss"]) - float(lo["loss"])) / abs(float(lo["loss"]))
    e["d F_color.0"] = rel_err(model.F_color[0].weight.grad, P.F_color[0][0].grad)
    e["d R.0"] = rel_err(model.R[0].weight.grad, P.R[0][0].grad)
    e["d latent_g"] = rel_err(model.neural_feats_geometry.grad, P.neural_feats_geometry.grad)
    e["d latent_c"] = rel_err(model.neural_feats_color.grad, P.neural_feats_color.grad)
    print("product fp32 mode vs reference GPU path (same sample depths):", {k: f"{v:.1e}" for k, v in e.items()})
    assert max(v for k, v in e.items() if not k.startswith("d ")) < 1e-4 and max(v for k, v in e.items() if k.startswith("d ")) < 1e-3, e


def test_sampler_device_noise_is_the_references_own():
    """"the eval sampler's tolerance is explained by cumsum noise -- asserted, not demonstrated".
    Demonstration: the reference's OWN sampler graph (oracle/hotpath.py::sample_z, pinned to the reference at 1e-5 on
    CPU) is run on the GPU, where torch's cumsum / sum are parallel reductions, and compared with the golden z the
    reference produced on CPU.  Same code, same inputs, other device: the inverse CDF divides by increments as small as
    1e-5 and the 5-iteration eval schedule is a chain of root searches, so the reference differs from ITSELF at the level
    the product differs from it.  The product is required to be no further from the golden than 3x the reference's own
    GPU-vs-CPU distance (bulk statistics), train and eval schedules."""
    from oracle import gpu_ref as G
    from spurfies_b200.model import PointVolSDF, default_conf
    from tests.helpers import load_golden, load_into_model
    if G.load_reference_ext() is None:
        pytest.skip("oracle/_ref/knnquery_cuda*.so not built")
    g, P = load_golden()
    model = load_into_model(PointVolSDF(default_conf(), "24", "dtu", neural_points=g["scene"]["pts"],
                                        neural_colors=g["scene"]["colors"]), P)
    _, P2 = load_golden()
    Pg = G.params_to(P2, "cuda")
    grid = G.make_grid("reference", Pg)
    cam = {"pose": g["pose"].cuda(), "intrinsics": g["intrinsics"].cuda()}
    uv = g["uv"].cuda()
    R = uv.shape[1]
    dirs, camloc = g["ray_dirs"].cuda(), g["cam_loc"].cuda().expand(R, 3).contiguous()
    rng = {k: v.cuda() for k, v in g["rng"].items()}
    stats = {}
    for name, training, fast, gold in (("train", True, 1, g["z_train"]), ("eval", False, -1, g["z_eval"])):
        with torch.no_grad():
            z_ref_gpu = G.sample_z(Pg, grid, uv, cam, training, fast, rng if training else None)
            model.train(training)
            z_prod, _ = model.ray_sampler.get_z_vals(dirs, camloc, model, fast, 1, rng=rng if training else None)
        for who, z in (("reference graph on GPU", z_ref_gpu), ("product", z_prod)):
            d = (z.cpu() - gold).abs()
            per_ray = d.max(-1).values
            stats[(name, who)] = {"median": float(d.median()), "ray p90": float(torch.quantile(per_ray, 0.9)), "max": float(d.max())}
            print(f"sampler {name:5s} vs CPU golden, {who:23s}:", {k: f"{v:.1e}" for k, v in stats[(name, who)].items()})
    for name in ("train", "eval"):
        a, b = stats[(name, "product")], stats[(name, "reference graph on GPU")]
        assert a["median"] <= 3 * b["median"] + 1e-6 and a["ray p90"] <= 3 * b["ray p90"] + 1e-5, (name, a, b)
