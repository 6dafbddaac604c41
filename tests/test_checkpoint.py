"""CPU: checkpoint compatibility with the reference (spurfies/train.py:123-154, 222-241, 292-328).  The names, shapes
and dtypes of the product model's state dict are compared with the reference's own ``PointVolSDF`` class
(tests/golden/checkpoint_spec.json, written by tests/golden/make_golden_checkpoint.py from /root/reference); the
reference-format checkpoint directory round-trips; the local-prior key mapping is the reference's."""
import json
import os
from collections import OrderedDict

import pytest
import torch

from spurfies_b200 import checkpoint as ck
from spurfies_b200 import scenes
from spurfies_b200.model import PointVolSDF, default_conf

SPEC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "checkpoint_spec.json")


def _model(n, seed=24):
    sc = scenes.dtu_like(n, seed=seed, radii=(0.3, 0.45))
    return PointVolSDF(default_conf(), "24", "dtu", neural_points=sc["pts"], neural_colors=sc["colors"], device="cpu")


@pytest.fixture(scope="module")
def spec():
    return json.load(open(SPEC))


def test_state_dict_has_the_reference_names_shapes_and_order(spec):
    m = _model(spec["n_points"])
    ours = [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()]
    assert ours == spec["state_dict"]
    # the order of named_parameters() is the order of torch.optim.Adam's parameter ids in a reference optimiser file
    assert [[k, list(p.shape)] for k, p in m.named_parameters()] == spec["named_parameters"]


def test_a_reference_checkpoint_loads_strictly(spec, tmp_path):
    # a state dict with exactly the reference's entries (random values), saved the way train.py:292-300 saves it
    g = torch.Generator().manual_seed(0)
    ref_sd = OrderedDict((k, torch.randn(shape, generator=g).to(getattr(torch, dt.split(".")[1])))
                         for k, shape, dt in spec["state_dict"])
    d = tmp_path / "checkpoints"
    os.makedirs(d / ck.MODEL_SUBDIR)
    torch.save({"epoch": 7, "model_state_dict": ref_sd, "iter_step": 1234}, d / ck.MODEL_SUBDIR / "latest.pth")
    m = _model(spec["n_points"])
    info = ck.load_from_dir(str(d), m)
    assert info == {"epoch": 7, "iter_step": 1234}
    for k, v in m.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
    # a checkpoint of another point count does not load silently
    with pytest.raises(RuntimeError):
        ck.load_from_dir(str(d), _model(spec["n_points"] + 1))


def test_checkpoint_directory_roundtrip(tmp_path):
    a, b = _model(300, seed=1), _model(300, seed=2)
    with torch.no_grad():
        a.density.beta.fill_(0.037)
    opt = torch.optim.Adam([p for p in a.parameters()], lr=5e-4)      # any object with a state_dict(): FusedAdam's has this layout
    sum((p ** 2).sum() for p in a.parameters()).backward()
    opt.step()
    d = str(tmp_path / "checkpoints")
    ck.save_checkpoints(d, 3, a, opt, iter_step=42, reference_layout=False)
    assert sorted(os.listdir(os.path.join(d, ck.MODEL_SUBDIR))) == ["3.pth", "latest.pth"]
    assert sorted(os.listdir(os.path.join(d, ck.OPTIM_SUBDIR))) == ["3.pth", "latest.pth"]
    raw = torch.load(os.path.join(d, ck.MODEL_SUBDIR, "3.pth"), weights_only=False)
    assert set(raw) == {"epoch", "model_state_dict", "iter_step"}                       # train.py:294-298
    assert set(torch.load(os.path.join(d, ck.OPTIM_SUBDIR, "3.pth"), weights_only=False)) == {"epoch", "optimizer_state_dict"}
    opt_b = torch.optim.Adam([p for p in b.parameters()], lr=1e-3)
    info = ck.load_from_dir(d, b, opt_b, checkpoint=3)
    assert info == {"epoch": 3, "iter_step": 42}
    for (k, va), (_, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(va, vb), k
    sa, sb = opt.state_dict(), opt_b.state_dict()
    assert sb["param_groups"][0]["lr"] == 5e-4
    for i in sa["state"]:
        assert torch.equal(sa["state"][i]["exp_avg"], sb["state"][i]["exp_avg"])
    ck.save_checkpoints(d, 4, a, opt, iter_step=43, latest_only=True)
    assert sorted(os.listdir(os.path.join(d, ck.MODEL_SUBDIR))) == ["3.pth", "latest.pth"]
    assert ck.load_from_dir(d, b)["iter_step"] == 43


def _fake_prior():
    """A file shaped like ckpt/local_prior.pt as train.py:125-139 consumes it: ``sdf_features`` first, then the five
    linears of the prior's SDF field (weight, bias each; four dotted components before the parameter name), then the
    density branch."""
    g = torch.Generator().manual_seed(3)
    sd = OrderedDict()
    sd["sdf_features"] = torch.randn(10, 32, generator=g)
    dims = [(256, 35), (256, 256), (256, 256), (256, 256), (256, 256)]
    for j, (o, i) in enumerate(dims):
        sd[f"model.field.local_sdf_field.{2 * j}.weight"] = torch.randn(o, i, generator=g)
        sd[f"model.field.local_sdf_field.{2 * j}.bias"] = torch.randn(o, generator=g)
    sd["model.field.density_branch.weight"] = torch.randn(1, 256, generator=g)
    sd["model.field.density_branch.bias"] = torch.randn(1, generator=g)
    return sd


def test_local_prior_mapping_and_freeze(tmp_path):
    prior = _fake_prior()
    mapped = ck.prior_to_model_state(prior)
    assert sorted(mapped) == sorted([f"F_geometry.{l}.{p}" for l in (0, 2, 4, 6, 8) for p in ("weight", "bias")]
                                    + ["T.0.weight", "T.0.bias"])
    assert mapped["F_geometry.4.bias"] is prior["model.field.local_sdf_field.4.bias"]
    m = _model(200)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    path = str(tmp_path / "local_prior.pt")
    torch.save({"model_state_dict": prior}, path)
    ck.load_prior(m, path)
    for k, v in m.state_dict().items():
        if k in mapped:
            assert torch.equal(v, mapped[k]), k
        else:
            assert torch.equal(v, before[k]), k                                          # strict=False: nothing else moves
    frozen = {n for n, p in m.named_parameters() if not p.requires_grad}
    assert frozen == set(mapped)                                                        # train.py:148-154
    bad = _fake_prior()
    bad["model.field.local_sdf_field.0.weight"] = torch.zeros(256, 36)
    with pytest.raises(ValueError):
        ck.load_prior(_model(200), bad)


def test_optimizer_file_loads_into_the_references_two_group_adam(tmp_path):
    """train.py:168-189 builds Adam([{"params": [] (sdf_feat), "lr": 1e-2}, {"params": trainable, "lr": lr}]); torch
    refuses a state dict with another group structure.  FusedAdam.state_dict() has ONE group: save_checkpoints writes
    the reference's layout around it."""
    m = _model(100)
    ck.load_prior(m, _fake_prior())
    params = [p for p in m.parameters() if p.requires_grad]
    n = len(params)
    g = torch.Generator().manual_seed(1)
    fused_like = {"state": {i: {"step": torch.tensor(5.0), "exp_avg": torch.randn(p.shape, generator=g),
                                "exp_avg_sq": torch.rand(p.shape, generator=g)} for i, p in enumerate(params)},
                  "param_groups": [{"lr": 4e-4, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 0, "amsgrad": False,
                                    "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                                    "fused": None, "params": list(range(n))}]}          # optim.py::FusedAdam.state_dict

    class Opt:
        def state_dict(self):
            return fused_like
    d = str(tmp_path / "checkpoints")
    ck.save_checkpoints(d, 1, m, Opt(), iter_step=5, latest_only=True)
    ref_opt = torch.optim.Adam([{"params": [], "lr": 1e-2}, {"params": params, "lr": 5e-4}])
    with pytest.raises(ValueError):
        ref_opt.load_state_dict(fused_like)
    ref_opt.load_state_dict(torch.load(os.path.join(d, ck.OPTIM_SUBDIR, "latest.pth"), weights_only=False)["optimizer_state_dict"])
    sd = ref_opt.state_dict()
    assert [len(gr["params"]) for gr in sd["param_groups"]] == [0, n] and sd["param_groups"][1]["lr"] == 4e-4
    assert all(torch.equal(sd["state"][i]["exp_avg"], fused_like["state"][i]["exp_avg"]) for i in range(n))
    assert ck.reference_optimizer_layout(sd) is sd          # already two groups: untouched
