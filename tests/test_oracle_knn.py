"""CPU: pin the kNN oracle (oracle/knn_oracle.c) against the reference's only known-answer test
(torch_knnquery/test/test_queries.py: seed 1234, cdist + topk) and against brute force on DTU-shaped scenes."""
import torch

from oracle.knn import OracleGrid, RefVoxelGrid, brute_force_neighbor_sets
from spurfies_b200 import scenes


def reference_test_inputs():
    """test_queries.py:11-18, 104-130 -- the draws come from the CPU generator there too (then .cuda())."""
    torch.manual_seed(1234)
    pts = 30 * (torch.rand(1, 1000, 3) - 0.5)
    rays_o = 3 * (torch.rand(1, 3, 100, 3) - 0.5)
    rays_d = torch.nn.functional.normalize(torch.rand(1, 3, 100, 3) - 0.5, dim=-1)
    depth = torch.linspace(-5, 5, 100)[None, None, :, None]
    return pts, rays_o + depth * rays_d
GRID_ARGS = dict(voxel_size=(1., 1., 1.), voxel_scale=(4., 4., 4.), kernel_size=(3, 3, 3),
                 ranges=(-20.0, -20.0, -20.0, 20.0, 20.0, 20.0))


def test_reference_kat_neighbor_sets():
    pts, raypos = reference_test_inputs()
    g = OracleGrid(pts, **GRID_ARGS)
    out = g.query_dense(raypos[0], 3, 1.0, 100)  # Smax >= D so that every mask-hit sample owns a slot
    want = brute_force_neighbor_sets(raypos[0], pts[0], 3, 1.0)  # [3,100,3] sorted, -1 padded
    got = torch.full((3, 100, 3), -1, dtype=torch.long)
    for r in range(3):
        for s in range(100):
            d = int(out["slot_sample"][r, s])
            if d >= 0:
                got[r, d] = out["pidx"][r, s].long().sort().values
    # a sample with a neighbour within r=1 always lies in a dilated-occupied voxel (voxel edge 4 > r)
    assert torch.equal(got, want)
    assert int((want >= 0).any(-1).sum()) > 20


def test_reference_kat_slot_rule_is_the_kernels():
    """With the test's own Smax=3 the kernel semantics (first 3 *mask-hit* samples) apply."""
    pts, raypos = reference_test_inputs()
    vg = RefVoxelGrid(max_points_per_voxel=1000, max_occ_voxels_per_example=610000, **GRID_ARGS)
    vg.set_pointset(pts, torch.tensor([1000], dtype=torch.int32))
    pidx, loc, ray_mask = vg.query(raypos, 3, 1.0, 3)
    m = vg.grid.mask(raypos[0])
    for r in range(3):
        first3 = torch.nonzero(m[r]).flatten()[:3]
        dense = vg.grid.query_dense(raypos[0], 3, 1.0, 3)
        assert torch.equal(dense["slot_sample"][r, :len(first3)].long(), first3)
    assert pidx.shape[1:] == (3, 3) and ray_mask.shape == (1, 3) and int(ray_mask.sum()) == pidx.shape[0]


def test_grid_equals_brute_force_on_dtu_scene():
    sc = scenes.dtu_like(20000, seed=3)
    g = OracleGrid(sc["pts"], (0.025,) * 3, (3,) * 3, (3,) * 3, sc["ranges"])
    gen = torch.Generator().manual_seed(0)
    q = sc["pts"][torch.randperm(20000, generator=gen)[:1500]] + 0.03 * torch.randn(1500, 3, generator=gen)
    grid = g.query_dense(q[:, None, :], 8, 2.0, 1)["pidx"][:, 0]
    brute = g.brute(q, 8, 2.0)
    hit = g.mask(q).bool()
    assert torch.equal(grid[hit], brute[hit])     # radius 0.05 < voxel 0.075: the 27 voxels hold every candidate
    assert int((brute[~hit] >= 0).sum()) == 0      # ... and a masked-out query has no neighbour within the radius
    want = brute_force_neighbor_sets(q, sc["pts"], 8, 0.05)
    same = (brute.long().sort(-1).values == want).all(-1)
    assert float(same.float().mean()) > 0.995      # cdist's sqrt/rounding may flip exact-radius / exact-tie cases


def test_edge_cases():
    pts = torch.tensor([[[0.0, 0.0, 0.0], [0.01, 0.0, 0.0], [0.5, 0.5, 0.5]]])
    g = OracleGrid(pts, (0.025,) * 3, (3,) * 3, (3,) * 3, (-1, -1, -1, 1, 1, 1))
    far = torch.full((2, 5, 3), 0.9)
    out = g.query_dense(far, 8, 2.0, 4)
    assert int(out["ray_mask1"].sum()) == 0 and int((out["pidx"] >= 0).sum()) == 0
    near = torch.zeros(1, 1, 3)
    out = g.query_dense(near, 8, 2.0, 1)
    assert out["pidx"][0, 0].tolist() == [0, 1, -1, -1, -1, -1, -1, -1]  # sorted by distance, -1 padded
    out = g.query_dense(near, 1, 2.0, 1)
    assert out["pidx"][0, 0].tolist() == [0]
    st = g.stats()
    assert st["occupied_voxels"] == 2 and st["max_points_per_voxel"] == 2 and st["points_in_grid"] == 3


def test_exact_ties_are_broken_by_point_id():
    """Duplicate and equidistant points: the K nearest are chosen and ordered by (d^2, point id) -- the deterministic rule
    the product kernels reproduce bit-exactly (the reference's own order depends on its atomics)."""
    c, h = 0.125, 1.0 / 64.0   # exactly representable: ids 5, 6, 7 are equidistant from the query in fp32 arithmetic
    dup = [[c, c, c]] * 5 + [[c + h, c, c], [c - h, c, c], [c, c + h, c]]
    pts = torch.tensor([dup])
    g = OracleGrid(pts, (0.025,) * 3, (3,) * 3, (3,) * 3, (-1, -1, -1, 1, 1, 1))
    q = torch.tensor([[[c, c, c]]])
    assert g.query_dense(q, 3, 2.0, 1)["pidx"][0, 0].tolist() == [0, 1, 2]
    assert g.query_dense(q, 7, 2.0, 1)["pidx"][0, 0].tolist() == [0, 1, 2, 3, 4, 5, 6]
    assert g.query_dense(q, 8, 2.0, 1)["pidx"][0, 0].tolist() == [0, 1, 2, 3, 4, 5, 6, 7]


def test_randomised_scenes_grid_equals_brute_force_and_slot_rule():
    """Property check over clustered random clouds, ray bundles and (K, Smax, D): every slot is one of the first Smax
    mask-hit samples of its ray, in order (knnquery.py:208-231); its neighbour list is the brute-force (d^2, id) top-K
    of that sample; rays without a hit are all -1; points outside `ranges` never appear (knnquery.cu:49)."""
    for seed, n, k, smax, d, rays in ((0, 3000, 8, 6, 24, 40), (1, 500, 1, 3, 16, 30), (2, 8000, 20, 24, 48, 25),
                                      (3, 50, 3, 80, 12, 20), (4, 4000, 8, 1, 1, 200)):
        gen = torch.Generator().manual_seed(seed)
        centres = (torch.rand(6, 3, generator=gen) - 0.5) * 1.4
        pts = centres[torch.randint(6, (n,), generator=gen)] + 0.06 * torch.randn(n, 3, generator=gen)
        pts[: n // 50] = (torch.rand(n // 50, 3, generator=gen) - 0.5) * 3.0          # some points outside the ranges
        ranges = (-1.0, -1.0, -1.0, 1.0, 1.0, 1.0)
        g = OracleGrid(pts[None], (0.025,) * 3, (3,) * 3, (3,) * 3, ranges)
        o = (torch.rand(rays, 1, 3, generator=gen) - 0.5) * 2.4
        tgt = centres[torch.randint(6, (rays,), generator=gen)][:, None] + 0.05 * torch.randn(rays, 1, 3, generator=gen)
        t = torch.linspace(0.0, 1.6, d)[None, :, None]
        raypos = (o + (tgt - o) * t).contiguous() if d > 1 else tgt.contiguous()
        out = g.query_dense(raypos, k, 2.0, smax)
        m = g.mask(raypos).bool()
        inside = ((pts > -1.0) & (pts < 1.0)).all(-1)
        for r in range(rays):
            hits = torch.nonzero(m[r]).flatten()[:smax]
            ss = out["slot_sample"][r]
            assert torch.equal(ss[:len(hits)].long(), hits) and bool((ss[len(hits):] < 0).all()), (seed, r)
            assert bool((out["pidx"][r, len(hits):] < 0).all())
            if len(hits):
                want = g.brute(raypos[r, hits], k, 2.0)
                assert torch.equal(out["pidx"][r, :len(hits)], want), (seed, r)
                assert torch.equal(out["sample_loc"][r, :len(hits)], raypos[r, hits])
        used = out["pidx"][out["pidx"] >= 0].long()
        assert bool(inside[used].all()), seed
        assert torch.equal(out["ray_mask1"].bool(), m.any(-1))
        assert torch.equal(out["ray_mask2"].bool(), (out["pidx"] >= 0).any(-1).any(-1))
        assert int((out["pidx"] >= 0).sum()) > 0 or n < 100
