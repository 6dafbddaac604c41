"""CPU: pin oracle/ingest.py (restatement of spurfies/model/utils.py:6-59) against golden vectors produced by the
reference's own functions (tests/golden/make_golden_ingest.py -> tests/golden/ingest.pt)."""
import os

import pytest
import torch

from oracle import ingest as OI

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ingest.pt")


def _cases():
    from spurfies_b200 import scenes
    for c in torch.load(GOLD, weights_only=False)["cases"]:
        pts = scenes.dtu_like(c["n"], seed=c["seed"], radii=tuple(c["radii"]))["pts"]
        assert abs(float(pts.double().abs().sum()) - c["pts_checksum"]) <= 1e-9 * c["pts_checksum"]
        yield c, pts


def test_restatement_matches_reference_voxelisation():
    n = 0
    for c, pts in _cases():
        cen, gidx, midx, res, inv = OI.construct_vox_points_closest(pts, c["vox_res"])
        assert torch.equal(gidx, c["grid_idx"])                       # same voxels, same (sorted) order
        assert float((cen - c["centroid"]).abs().max()) < 2e-7        # fp64-summed vs the reference's fp32-summed mean
        same = midx == c["min_idx"]
        # the kept point can differ only where two points are equidistant from the centroid to within that ulp
        assert float(same.float().mean()) > 0.999
        d = (~same).nonzero().flatten()
        if len(d):
            assert torch.equal(inv[c["min_idx"][d]], d)
            assert float((res[c["min_idx"][d]] - res[midx[d]]).abs().max()) < 1e-6
        kept, idx = OI.voxelize(pts, c["vox_res"])
        assert torch.equal(idx, midx) and torch.equal(kept, pts[idx])
        n += 1
    assert n == 3
