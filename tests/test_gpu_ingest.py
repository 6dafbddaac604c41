"""GPU: SURVEY 8(f3) -- neural-point ingestion (PLY -> voxel down-sampling -> model) against the CPU restatement of
spurfies/model/utils.py:6-88.  Voxel sets, order and integer coordinates: exact.  Kept point per voxel: identical except
where two points are equidistant from the centroid to within fp32 rounding (the reference itself is not reproducible
there: its scatter_mean / scatter_min run on float atomics)."""
import numpy as np
import pytest
import torch

from oracle import ingest as OI

pytestmark = pytest.mark.gpu


def _cloud(n, seed):
    from spurfies_b200 import scenes
    sc = scenes.dtu_like(n, seed=seed, radii=(0.35, 0.5))
    return sc["pts"], sc["colors"]


@pytest.mark.parametrize("n,vox_res", [(20000, 50), (200000, 300), (1000, 7)])
def test_voxel_downsample_matches_restated_reference(n, vox_res):
    from spurfies_b200 import ingest
    pts, _ = _cloud(n, 3)
    cen, gidx, midx = ingest.construct_vox_points_closest(pts.cuda(), vox_res)
    rc, rg, rm, res, inv = OI.construct_vox_points_closest(pts, vox_res)
    assert torch.equal(gidx.cpu(), rg)                         # same voxels, same (sorted) order
    assert float((cen.cpu() - rc).abs().max()) < 1e-6
    got = midx.cpu()
    same = got == rm
    assert float(same.float().mean()) > 0.999
    # where the pick differs it is a numerical tie: same voxel, residual within rounding of the minimum
    d = (~same).nonzero().flatten()
    assert torch.equal(inv[got[d]], d)
    assert float((res[got[d]] - res[rm[d]]).abs().max() if len(d) else 0.0) < 1e-6
    assert len(torch.unique(got)) == len(got)


def test_ply_roundtrip_and_model_from_ply(tmp_path):
    from spurfies_b200 import ingest
    from spurfies_b200.model import PointVolSDF, default_conf
    pts, colors = _cloud(30000, 5)
    path = str(tmp_path / "24.ply")
    ingest.write_ply(path, pts.numpy(), colors.numpy())
    ply = ingest.read_ply(path)
    assert np.array_equal(ply["x"], pts[:, 0].numpy()) and np.array_equal(ply["blue"], colors[:, 2].numpy().astype(np.uint8))
    data = ingest.load_neural_points(path, vox_res=60)
    want_pts, want_idx = OI.voxelize(pts, 60)
    assert data["pts"].shape == want_pts.shape and data["colors"].shape == want_pts.shape
    same = (data["pts"].cpu() == want_pts).all(dim=1)
    assert float(same.float().mean()) > 0.99
    # a different kept point is a numerical tie (see the module docstring): its distance to the voxel centroid is
    # within fp32 rounding of the minimum the restatement found
    _, _, rm, res, inv = OI.construct_vox_points_closest(pts, 60)
    d = (~same).nonzero().flatten()
    got = (pts[None, :, :] == data["pts"].cpu()[d][:, None, :]).all(dim=2).float().argmax(dim=1)
    assert torch.equal(inv[got], d) and float((res[got] - res[rm[d]]).abs().max() if len(d) else 0.0) < 1e-6
    # ascii PLY without colours
    apath = str(tmp_path / "a.ply")
    with open(apath, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment test\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                "end_header\n0 0 0\n1 2 3\n-1 0.5 2\n")
    a = ingest.load_neural_points(apath)
    assert "colors" not in a and a["pts"].cpu().tolist() == [[0, 0, 0], [1, 2, 3], [-1, 0.5, 2]]
    # the model constructor takes the reference's route: conf.pointcloud_path + conf.vox_res (pointneus_disent.py:131-146)
    conf = default_conf(pointcloud_path=path, vox_res=60)
    model = PointVolSDF(conf, "24", "dtu")
    assert model.neural_pts.shape == want_pts.shape and model.neural_pts.is_cuda
    c = model.neural_feats_color[:, :3].detach().cpu()
    assert float((c - (data["colors"].cpu().float() * 2.0 / 255.0 - 1.0)).abs().max()) < 1e-6
    with torch.no_grad():
        s = model.get_sdf_eval(model.neural_pts[:64].contiguous())
    assert bool((s != 1000).all())
    with pytest.raises(RuntimeError):
        PointVolSDF(default_conf(pointcloud_path=str(tmp_path / "missing.ply")), "24", "dtu")


def test_voxel_downsample_matches_reference_golden():
    """The kernels against tests/golden/ingest.pt -- outputs of the REFERENCE's own construct_vox_points_closest /
    voxelize (spurfies/model/utils.py:6-59, imported with a pure-torch torch_scatter shim by
    tests/golden/make_golden_ingest.py): voxel set and order exact, centroid within fp32 rounding, kept point per voxel
    identical except at numerical ties."""
    import os
    from spurfies_b200 import ingest, scenes
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ingest.pt"), weights_only=False)
    for c in gold["cases"]:
        pts = scenes.dtu_like(c["n"], seed=c["seed"], radii=tuple(c["radii"]))["pts"]
        assert abs(float(pts.double().abs().sum()) - c["pts_checksum"]) <= 1e-9 * c["pts_checksum"]
        cen, gidx, midx = ingest.construct_vox_points_closest(pts.cuda(), c["vox_res"])
        assert torch.equal(gidx.cpu(), c["grid_idx"])
        assert float((cen.cpu() - c["centroid"]).abs().max()) < 1e-6
        got, ref = midx.cpu(), c["min_idx"]
        same = got == ref
        assert float(same.float().mean()) > 0.995, float(same.float().mean())
        d = (~same).nonzero().flatten()
        if len(d):   # every difference must be a numerical tie: distance to the centroid within rounding of the reference's minimum
            r_got = (pts[got[d]] - c["centroid"][d]).norm(dim=-1)
            r_ref = (pts[ref[d]] - c["centroid"][d]).norm(dim=-1)
            assert float((r_got - r_ref).abs().max()) < 1e-6
