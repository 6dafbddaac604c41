"""Golden vectors for the camera-file readers from the REFERENCE's own
``spurfies/utils/rend_util.py::load_K_Rt_from_P`` (:36-57, which calls cv2.decomposeProjectionMatrix), imported from
/root/reference with its I/O-only imports (imageio, skimage) stubbed.  The output file IS a DTU-format ``cameras.npz``
(``world_mat_i`` / ``scale_mat_i`` as spurfies/datasets/dtu.py:78-86 reads them) plus the reference's answers
``ref_intrinsics_i`` / ``ref_pose_i`` for ``P = (world_mat @ scale_mat)[:3, :4]``.  Runs only in the authoring
container; the tests use the committed tests/golden/cameras.npz.

Usage:  python tests/golden/make_golden_cameras.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("SPF_REFERENCE_ROOT", "/root/reference")


def synthetic_dtu_cameras(n=8, seed=24):
    """world_mat = K4 @ [R | t] (world -> pixel, DTU convention), scale_mat = the normalisation that maps the unit
    sphere onto the object (uniform scale + translation)."""
    from spurfies_b200 import scenes
    rng = np.random.default_rng(seed)
    out = {}
    centre, radius = np.array([12.0, -30.0, 610.0]), 180.0
    for i in range(n):
        eye = centre + radius * 2.3 * scenes._unit(rng.normal(size=3))
        c2w = scenes.look_at_pose(eye, target=centre + rng.normal(size=3) * 5.0)
        w2c = np.linalg.inv(c2w)
        K = np.eye(4)
        K[0, 0], K[1, 1] = 2892.33 * (1 + 0.01 * rng.normal()), 2883.18 * (1 + 0.01 * rng.normal())
        K[0, 1] = 0.3 * rng.normal()            # a little skew, as calibrated cameras have
        K[0, 2], K[1, 2] = 823.2 + rng.normal(), 619.07 + rng.normal()
        wm = K @ w2c
        if i == n - 1:
            wm = -wm                             # P and -P are the same camera: pins the sign convention
        sm = np.eye(4)
        sm[:3, :3] *= radius
        sm[:3, 3] = centre
        out["world_mat_%d" % i] = wm.astype(np.float64)
        out["scale_mat_%d" % i] = sm.astype(np.float64)
    return out, n


def main():
    for name in ("imageio", "skimage", "skimage.transform"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage"].img_as_float32 = None
    sys.modules["skimage.transform"].rescale = None
    spec = importlib.util.spec_from_file_location("ref_rend_util", os.path.join(REF, "spurfies", "utils", "rend_util.py"))
    RU = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(RU)
    cams, n = synthetic_dtu_cameras()
    for i in range(n):
        sm, wm = cams["scale_mat_%d" % i].astype(np.float32), cams["world_mat_%d" % i].astype(np.float32)   # dtu.py:79-86
        P = (wm @ sm)[:3, :4]
        intr, pose = RU.load_K_Rt_from_P(None, P)                                                               # dtu.py:113
        cams["ref_intrinsics_%d" % i], cams["ref_pose_%d" % i] = intr, pose
    path = os.path.join(ROOT, "tests", "golden", "cameras.npz")
    np.savez(path, **cams)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
