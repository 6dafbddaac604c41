"""CPU: the oracle restatement of the feature-consistency loss (oracle/local_loss.py) against golden vectors produced
by the reference's own feat_utils.get_local_loss / find_surface_points (tests/golden/make_golden_local.py)."""
import os

import pytest
import torch

from oracle import local_loss as LL
from spurfies_b200 import scenes

GOLD = os.path.join(os.path.dirname(__file__), "golden", "local_loss.pt")
TOL = 1e-5


@pytest.fixture(scope="module")
def gold():
    g = torch.load(GOLD)
    ld = scenes.local_data(0, 2.3, feat_res=tuple(g["feat_res"]))
    for k, v in g["local_data_checksum"].items():   # the regenerated synthetic features are the ones the reference saw
        assert abs(float(ld[k].double().abs().sum()) - v) <= 1e-6 * max(1.0, abs(v)), k
    return g, ld


def test_find_surface_points_matches_reference(gold):
    g, _ = gold
    d, m = LL.find_surface_points(g["sdf"], g["z"])
    assert torch.equal(m, g["network_mask"])
    assert torch.allclose(d, g["d_surface"], rtol=TOL, atol=1e-6)
    assert g["empty_mask_sum"] == 0
    d2, m2 = LL.find_surface_points(torch.full((4, 80), 1000.0), torch.zeros(4, 80))
    assert int(m2.sum()) == 0 and float(d2.abs().sum()) == 0.0


def test_local_loss_value_and_gradient_match_reference(gold):
    g, ld = gold
    sdf = g["sdf"].clone().requires_grad_(True)
    loss = LL.local_loss_from_rays(sdf, g["z"], g["cam_loc"], g["ray_dirs"], ld)
    assert abs(float(loss) - float(g["loss"])) <= TOL * max(1.0, abs(float(g["loss"])))
    loss.backward()
    gr = torch.nan_to_num(sdf.grad, nan=0.0)
    err = (gr - g["d_sdf"]).abs().max() / g["d_sdf"].abs().max()
    assert err < 1e-4, float(err)
    # the gradient touches exactly the two slots either side of each crossing
    assert int((gr != 0).sum()) <= 2 * int(g["network_mask"].sum())


def test_local_loss_empty_is_zero(gold):
    _, ld = gold
    z = torch.zeros(3, 80)
    assert float(LL.local_loss_from_rays(torch.full((3, 80), 1000.0), z, torch.zeros(3, 3), torch.ones(3, 3), ld)) == 0.0
