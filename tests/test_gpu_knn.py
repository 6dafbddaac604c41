"""GPU: the sm_100a voxel-grid kNN (through the C ABI / VoxelGrid drop-in) against the CPU oracle, the
reference's known-answer test, and -- when oracle/_ref holds it -- the unmodified reference CUDA extension."""
import glob
import importlib.util
import os
import time

import pytest
import torch

from oracle.knn import OracleGrid, brute_force_neighbor_sets
from spurfies_b200 import scenes
from tests.test_oracle_knn import GRID_ARGS, reference_test_inputs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_grid(pts, ranges, P=26, max_o=20000, vs=(0.025,) * 3, sc=(3,) * 3, ks=(3,) * 3):
    from spurfies_b200.knnquery import VoxelGrid
    vg = VoxelGrid(vs, sc, ks, P, max_o, ranges).cuda()
    p = pts.reshape(1, -1, 3).cuda().contiguous()
    vg.set_pointset(p, torch.tensor([p.shape[1]], dtype=torch.int32, device="cuda"))
    return vg


def test_reference_kat_through_dropin_api():
    """test_queries.py:101-153 re-expressed against our VoxelGrid: sorted index sets, sample locations, ray mask."""
    from spurfies_b200.knnquery import VoxelGrid
    pts, raypos = reference_test_inputs()
    vg = VoxelGrid(max_points_per_voxel=1000, max_occ_voxels_per_example=610000, **GRID_ARGS).cuda()
    vg.set_pointset(pts.cuda(), 1000 * torch.ones(1, dtype=torch.int).cuda())
    for smax in (100, 3):
        pidx, loc, ray_mask = vg.query(raypos.cuda().contiguous(), 3, 1.0, smax)
        o = OracleGrid(pts, **GRID_ARGS).query_dense(raypos[0], 3, 1.0, smax)
        keep = o["ray_mask2"].bool()
        assert pidx.dtype == torch.int32 and ray_mask.dtype == torch.int8 and tuple(ray_mask.shape) == (1, 3)
        assert torch.equal(ray_mask.cpu()[0].bool(), keep)
        assert torch.equal(pidx.cpu(), o["pidx"][keep])          # bit-exact, including the (d2, id) order
        assert torch.equal(loc.cpu(), o["sample_loc"][keep])
    # per-sample neighbour sets equal the reference test's brute-force cdist/topk oracle
    pd, _, slot_sample, _ = vg.query_dense(raypos[0].cuda().contiguous(), 3, 1.0, 100)
    want = brute_force_neighbor_sets(raypos[0], pts[0], 3, 1.0)
    got = torch.full((3, 100, 3), -1, dtype=torch.long)
    ss, pdc = slot_sample.cpu(), pd.cpu().long()
    for r in range(3):
        for s in range(100):
            if ss[r, s] >= 0:
                got[r, ss[r, s]] = pdc[r, s].sort().values
    assert torch.equal(got, want)


@pytest.mark.parametrize("n_points,k", [(10000, 8), (100000, 8), (100000, 20), (30000, 1)])
def test_rays_match_oracle_bit_exact(n_points, k):
    sc = scenes.dtu_like(n_points, seed=24)
    vg = make_grid(sc["pts"], sc["ranges"])
    og = OracleGrid(sc["pts"], (0.025,) * 3, (3,) * 3, (3,) * 3, sc["ranges"])
    assert vg.stats() == og.stats()
    from oracle.hotpath import camera_rays
    cam = scenes.camera(1, sc["cam_radius"])
    uv = scenes.pixel_batch(256, seed=2)
    d, o = camera_rays(uv, cam["pose"], cam["intrinsics"])
    z = torch.sort(torch.rand(256, 98, generator=torch.Generator().manual_seed(1)) * 3.0 + 0.8, -1).values
    pts = (o[:, None, :] + z[..., None] * d[0][:, None, :]).contiguous()
    for smax in (80, 5):
        pidx, loc, slot_sample, nvalid = vg.query_dense(pts.cuda(), k, 2.0, smax)
        ref = og.query_dense(pts, k, 2.0, smax)
        assert torch.equal(slot_sample.cpu(), ref["slot_sample"])
        assert torch.equal(loc.cpu(), ref["sample_loc"])
        assert torch.equal(pidx.cpu(), ref["pidx"])
        assert torch.equal((nvalid.cpu() > 0), ref["ray_mask2"].bool())
        assert int((ref["pidx"] >= 0).sum()) > 50


def test_points_match_oracle_and_mask():
    sc = scenes.dtu_like(100000, seed=24)
    vg = make_grid(sc["pts"], sc["ranges"])
    og = OracleGrid(sc["pts"], (0.025,) * 3, (3,) * 3, (3,) * 3, sc["ranges"])
    g = torch.Generator().manual_seed(5)
    q = torch.cat([sc["pts"][:20000] + 0.02 * torch.randn(20000, 3, generator=g), torch.rand(20001, 3, generator=g) * 2.4 - 1.2])
    pidx = vg.query_points(q.cuda().contiguous(), 8, 2.0)
    ref = og.query_dense(q[:, None, :], 8, 2.0, 1)
    assert torch.equal(pidx.cpu(), ref["pidx"][:, 0])
    assert torch.equal(vg.mask_points(q.cuda().contiguous()).cpu(), og.mask(q))
    # compaction
    from spurfies_b200.fields import SlotSet
    s = SlotSet(pidx)
    valid = torch.nonzero(ref["pidx"][:, 0, 0] >= 0).flatten()
    assert s.V == len(valid) and torch.equal(s.list[:s.V].cpu().long(), valid)


def test_edge_cases():
    pts = torch.tensor([[0.0, 0.0, 0.0], [0.01, 0.0, 0.0], [0.5, 0.5, 0.5]])
    vg = make_grid(pts, (-1, -1, -1, 1, 1, 1))
    far = torch.full((2, 5, 3), 0.9).cuda()
    pidx, loc, ray_mask = vg.query(far[None], 8, 2.0, 4)
    assert pidx.shape == (0, 4, 8) and loc.shape == (0, 4, 3) and int(ray_mask.sum()) == 0
    near = torch.zeros(1, 1, 1, 3).cuda()
    pidx, loc, ray_mask = vg.query(near, 8, 2.0, 1)
    assert pidx[0, 0].tolist() == [0, 1, -1, -1, -1, -1, -1, -1]
    with pytest.raises(AssertionError):
        vg.query(near, 21, 2.0, 1)  # knnquery.py:184
    with pytest.raises(Exception):
        vg.query(near.cpu(), 8, 2.0, 1)
    assert vg.query_points(torch.zeros(0, 3).cuda(), 8, 2.0).shape == (0, 8)
    assert not vg.caps_exceeded()
    # point set exceeding the reference caps: we keep every point and say so
    dense = torch.rand(5000, 3, generator=torch.Generator().manual_seed(0)) * 0.05
    vg2 = make_grid(dense, (-1, -1, -1, 1, 1, 1))
    assert vg2.caps_exceeded() and vg2.stats()["points_in_grid"] == 5000


def test_full_size_properties():
    """BASELINE config sizes (1 M points, 8192 x 98 samples): size-independent properties instead of the oracle:
    sorted by distance, within radius, ids unique, self-query returns self first, idempotent."""
    sc = scenes.garden_like(1_000_000)
    vg = make_grid(sc["pts"], sc["ranges"])
    P = sc["pts"].cuda()
    q = P[:300000].contiguous()
    pidx = vg.query_points(q, 8, 2.0)
    assert torch.equal(pidx, vg.query_points(q, 8, 2.0))
    valid = pidx >= 0
    nb = P[pidx.clamp(min=0).long()]
    d2 = ((nb - q[:, None, :]) ** 2).sum(-1)
    assert bool((d2[valid] <= 0.0025 * (1 + 1e-5)).all())
    d2m = torch.where(valid, d2, torch.full_like(d2, 1e9))
    assert bool((d2m[:, 1:] >= d2m[:, :-1]).all())
    assert bool((d2[:, 0] == 0).all())                     # a stored point finds itself (or a duplicate) at distance 0
    srt = torch.where(valid, pidx, torch.arange(-8, 0, device="cuda", dtype=torch.int32)[None]).sort(-1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())
    st = vg.stats()
    assert st["points_in_grid"] == 1_000_000


@pytest.mark.parametrize("scene,n_points", [("dtu", 100_000), ("garden", 1_000_000)])
def test_thread_per_query_kernels_equal_warp_per_query_kernels(scene, n_points):
    """spf_knn_set_algo: the thread-per-query kernels (K <= 8, the hot path) and the warp-per-query kernels select the
    same K smallest (d2, id) -- identical index lists in identical order -- for ray slots (ragged slot counts, rays that
    miss) and for point queries (masked-out points, empty neighbourhoods), at BASELINE configs[1] / configs[2] size."""
    from spurfies_b200 import _lib
    sc = scenes.dtu_like(n_points) if scene == "dtu" else scenes.garden_like(n_points)
    vg = make_grid(sc["pts"], sc["ranges"])
    g = torch.Generator().manual_seed(11)
    P = sc["pts"]
    near = (P[torch.randint(0, n_points, (200_000,), generator=g)] + 0.02 * torch.randn(200_000, 3, generator=g))
    far = (torch.rand(60_000, 3, generator=g) * 2 - 1) * float(sc["ranges"][3])
    q = torch.cat([near, far])[torch.randperm(260_000, generator=g)].cuda().contiguous()
    # rays: 2048 x 98 samples marching through the object (consecutive samples close together, like the sampler's)
    o = (torch.rand(2048, 1, 3, generator=g) * 2 - 1) * 0.9 * float(sc["ranges"][3])
    d = torch.nn.functional.normalize(torch.randn(2048, 1, 3, generator=g), dim=-1)
    t = torch.sort(torch.rand(2048, 98, 1, generator=g) * 1.2 - 0.6, dim=1).values
    rays = (o + d * t).cuda().contiguous()
    out, ms = {}, {}
    try:
        for algo in (1, 2):
            _lib.call("spf_knn_set_algo", algo)
            for k in (8, 3):
                res = []
                for rep in range(2):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e2 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                    a = vg.query_points(q, k, 2.0)
                    e2.record()
                    b = vg.query_dense(rays, k, 2.0, 80)
                    e1.record()
                    torch.cuda.synchronize()
                    res = (a, b[0], b[3])
                    ms[(algo, k)] = (round(e0.elapsed_time(e2), 3), round(e2.elapsed_time(e1), 3))
                out[(algo, k)] = res
    finally:
        _lib.call("spf_knn_set_algo", 0)
    # the kernel-family hint of the automatic mode: set from the build statistics (very dense voxels -> warp per query)
    assert vg.handle.dense_cloud == (1 if scene == "garden" else 0), vg.stats()
    print(f"{scene}: ms (points, slots) per kernel family (1 = warp per query, 2 = thread per query):", {f"algo{a}_k{k}": v for (a, k), v in ms.items()})
    for k in (8, 3):
        for x, y in zip(out[(2, k)], out[(1, k)]):
            assert torch.equal(x, y)
    assert int((out[(2, 8)][0] >= 0).sum()) > 500_000 and int((out[(2, 8)][0][:, 0] < 0).sum()) > 10_000
    assert int((out[(2, 8)][1] >= 0).sum()) > 100_000


def test_search_grid_equals_reference_voxel_scan():
    """The radius-sized search grid is an access-path optimisation only: identical indices (same order) as scanning the
    27 reference voxels, for ray queries and point queries; and it switches itself off when radius > voxel edge."""
    sc = scenes.dtu_like(60000, seed=5)
    q = (sc["pts"][:40000] + 0.02 * torch.randn(40000, 3, generator=torch.Generator().manual_seed(1))).cuda().contiguous()
    rays = q[:32768].view(512, 64, 3).contiguous()
    out = {}
    for use in (True, False):
        vg = make_grid(sc["pts"], sc["ranges"], P=128, max_o=32768)
        vg.use_search_grid = use
        out[use] = (vg.query_points(q, 8, 2.0), vg.query_dense(rays, 8, 2.0, 48)[0], vg.query_points(q, 8, 4.0))
        assert (vg.handle.search_sorted is not None and vg.handle.search_sorted != 0) == use
    for a, b in zip(out[True], out[False]):
        assert torch.equal(a, b)
    assert int((out[True][0] >= 0).sum()) > 100000


def _load_reference_ext():
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "knnquery_cuda*.so"))
    if not so:
        return None
    spec = importlib.util.spec_from_file_location("knnquery_cuda", so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_against_unmodified_reference_extension():
    """Drive the reference's own kernels (compiled unmodified by oracle/build_ref.sh) the way knnquery.py:52-285
    drives them, with caps above the occupancy so its answer is well defined, and compare index SETS."""
    ext = _load_reference_ext()
    if ext is None:
        pytest.skip("oracle/_ref/knnquery_cuda*.so not built (needs /root/reference at build time)")
    sc = scenes.dtu_like(100000, seed=24)
    vg = make_grid(sc["pts"], sc["ranges"], P=128, max_o=32768)
    dev = "cuda"
    points = sc["pts"].reshape(1, -1, 3).cuda().contiguous()
    B, N, P, max_o, K, Smax = 1, points.shape[1], 128, 32768, 8, 80
    g = vg.handle
    shift = torch.tensor(list(g.shift), dtype=torch.float32, device=dev)
    vsize = torch.tensor(list(g.vsize), dtype=torch.float32, device=dev)
    vdim = torch.tensor(list(g.dim), dtype=torch.int32, device=dev)
    ks = torch.tensor([3, 3, 3], dtype=torch.int32, device=dev)
    G = g.n_cells
    dims = list(g.dim)
    coor_occ = torch.zeros([B] + dims, dtype=torch.int32, device=dev)
    occ_2_pnts = torch.full([B, max_o, P], -1, dtype=torch.int32, device=dev)
    occ_2_coor = torch.full([B, max_o, 3], -1, dtype=torch.int32, device=dev)
    occ_numpnts = torch.zeros([B, max_o], dtype=torch.int32, device=dev)
    coor_2_occ = torch.full([B] + dims, -1, dtype=torch.int32, device=dev)
    occ_idx = torch.zeros([B], dtype=torch.int32, device=dev)
    n_t = torch.tensor([N], dtype=torch.int32, device=dev)
    sec = int(round(time.time() * 1000))
    ext.find_occupied_voxels(points, n_t, B, N, shift, vsize, vdim, G, max_o, occ_idx, coor_2_occ, occ_2_coor, sec)
    coor_2_occ = torch.full([B] + dims, -1, dtype=torch.int32, device=dev)
    ext.create_coor_occ_maps(B, vdim, ks, G, max_o, occ_idx, coor_occ, coor_2_occ, occ_2_coor)
    ext.assign_points_to_occ_voxels(points, n_t, B, N, P, shift, vsize, vdim, G, max_o, coor_2_occ, occ_2_pnts,
                                    occ_numpnts, sec)
    torch.cuda.synchronize()
    assert int(occ_idx[0]) == vg.stats()["occupied_voxels"] and int(occ_numpnts.max()) == vg.stats()["max_points_per_voxel"]
    assert torch.equal(coor_occ.view(-1).bool(), vg._hit.bool())
    from oracle.hotpath import camera_rays
    cam = scenes.camera(0, sc["cam_radius"])
    uv = scenes.pixel_batch(512, seed=3)
    d, o = camera_rays(uv, cam["pose"], cam["intrinsics"])
    z = torch.sort(torch.rand(512, 98, generator=torch.Generator().manual_seed(1)) * 3.0 + 0.8, -1).values
    raypos = (o[:, None, :] + z[..., None] * d[0][:, None, :]).contiguous().cuda()
    R, D = 512, 98
    mask = torch.zeros([B, R, D], dtype=torch.int32, device=dev)
    ext.create_raypos_mask(raypos[None].contiguous(), coor_occ, B, R, D, G, shift, vdim, vsize, mask)
    mask = mask.view(R, D)
    ray_mask_1 = mask.max(-1)[0] > 0
    R_valid = int(ray_mask_1.sum())
    rp = raypos[ray_mask_1].contiguous()
    m = mask[ray_mask_1]
    cum = torch.cumsum(m, dim=-1).to(torch.int32)
    m = (m * cum * (cum <= Smax) - 1).contiguous()
    loc = torch.zeros([R_valid, Smax, 3], dtype=torch.float32, device=dev)
    loc_mask = torch.zeros([R_valid, Smax], dtype=torch.int32, device=dev)
    ref_pidx = torch.full([R_valid, Smax, K], -1, dtype=torch.int32, device=dev)
    ext.get_shadingloc(rp, m, R_valid, D, Smax, loc, loc_mask)
    r2b = torch.zeros(R_valid, dtype=torch.int32, device=dev)
    ext.query_along_ray(points, r2b, R_valid, Smax, max_o, P, K, G, vg.radius2(2.0), shift, vdim, vsize, ks,
                        occ_numpnts, occ_2_pnts, coor_2_occ, loc, loc_mask, ref_pidx)
    torch.cuda.synchronize()
    pidx, myloc, slot_sample, nvalid = vg.query_dense(raypos, K, 2.0, Smax)
    assert torch.equal(myloc[ray_mask_1], loc)
    a = pidx[ray_mask_1].sort(-1).values
    b = ref_pidx.sort(-1).values
    assert torch.equal(a, b), f"{int((a != b).any(-1).sum())} of {a.shape[0] * a.shape[1]} slots differ"
    assert int((b >= 0).sum()) > 10000


def test_point_queries_under_a_device_side_predicate():
    """spf_knn_points_pred (VoxelGrid.query_points(skip=)): flag 0 -> the plain query; flag != 0 at launch time -> nothing
    is searched, every row is -1 (the eval sampler's iterations after convergence); both kernel families."""
    from spurfies_b200 import _lib
    sc = scenes.dtu_like(30000, seed=3)
    vg = make_grid(sc["pts"], sc["ranges"])
    q = (sc["pts"][:20000] + 0.01 * torch.randn(20000, 3, generator=torch.Generator().manual_seed(2))).cuda().contiguous()
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    try:
        for algo in (1, 2):
            _lib.call("spf_knn_set_algo", algo)
            plain = vg.query_points(q, 8, 2.0)
            assert int((plain >= 0).sum()) > 50000
            flag.zero_()
            assert torch.equal(vg.query_points(q, 8, 2.0, skip=flag), plain)
            flag.fill_(1)
            assert bool((vg.query_points(q, 8, 2.0, skip=flag) == -1).all())
    finally:
        _lib.call("spf_knn_set_algo", 0)
