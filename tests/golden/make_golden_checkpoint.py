"""Golden for checkpoint compatibility: the REFERENCE's own ``PointVolSDF`` class (spurfies/model/pointneus_disent.py,
imported from /root/reference through the shims of make_golden.py) is instantiated on a synthetic point set and the
names / shapes / dtypes of its ``state_dict()`` and of its ``named_parameters()`` (in registration order, which is
the order of ``torch.optim.Adam``'s parameter ids, train.py:168-189) are written to tests/golden/checkpoint_spec.json.
Runs only in the authoring container.

Usage:  python tests/golden/make_golden_checkpoint.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as MG  # noqa: E402


def main():
    MG.install_shims()
    from spurfies_b200 import scenes
    n = 1500
    scene = scenes.dtu_like(n, seed=24, radii=(0.3, 0.45))
    model = MG.build_reference_model(scene)
    spec = {"n_points": n,
            "state_dict": [[k, list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()],
            "named_parameters": [[k, list(p.shape)] for k, p in model.named_parameters()],
            "frozen_rule": "train.py:148-154: names containing 'F_geometry' or 'T.0' get requires_grad_(False)"}
    path = os.path.join(MG.ROOT, "tests", "golden", "checkpoint_spec.json")
    with open(path, "w") as f:
        json.dump(spec, f)
    print("wrote", path, len(spec["state_dict"]), "entries")


if __name__ == "__main__":
    main()
