"""Camera files on either side of the hot path ("format compatibility"): what the reference's datasets
hand to ``PointVolSDF.forward`` as ``intrinsics`` / ``pose``.  Host-side numpy, no GPU work: these run once per scene.

* DTU ``cameras.npz`` (``world_mat_i`` / ``scale_mat_i``): spurfies/datasets/dtu.py:78-86, 107-120 -- projection
  ``P = (world_mat @ scale_mat)[:3, :4]`` decomposed by ``rend_util.load_K_Rt_from_P`` (rend_util.py:36-57, which calls
  ``cv2.decomposeProjectionMatrix``) into a 4x4 intrinsics matrix and a camera-to-world pose; intrinsics rows scaled to
  the training resolution.
* camera JSON (``fl_x, fl_y, cx, cy, w, h, frames[].file_path / transform_matrix``): written by
  dust3r_inference_own.py:161-181 (``save_json``), read by spurfies/datasets/own_data.py:55-76 and
  mip_nerf.py:73-150 (which keeps a named subset of the frames).

OpenCV is not needed: the RQ decomposition is done here with the sign convention OpenCV documents for
``RQDecomp3x3`` (rotation with determinant +1, the first two diagonal entries of K positive), which makes it unique.
"""
from __future__ import annotations

import json
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


def decompose_projection(P: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """P [3,4] -> (K [3,3] upper triangular, R [3,3] rotation, c [3] camera centre) with ``P[:, :3] = K @ R`` and
    ``P @ [c, 1] = 0``: the three outputs of ``cv2.decomposeProjectionMatrix`` that rend_util.py:44-47 reads
    (its homogeneous ``t`` is ``[c, 1]`` up to scale).  float64 arithmetic."""
    P = np.asarray(P, dtype=np.float64)
    if P.shape != (3, 4):
        raise ValueError(f"decompose_projection: P must be [3,4], got {P.shape}")
    M = P[:, :3]
    if abs(np.linalg.det(M)) < 1e-300:
        raise ValueError("decompose_projection: the left 3x3 block is singular (camera at infinity)")
    # RQ through the QR of the row-reversed transpose
    J = np.flipud(np.eye(3))
    q, r = np.linalg.qr((J @ M).T)
    K = J @ r.T @ J
    R = J @ q.T
    # unique form: K[0,0] > 0, K[1,1] > 0, det R = +1 (the sign of K[2,2] then follows det M)
    d = np.array([1.0 if K[0, 0] >= 0 else -1.0, 1.0 if K[1, 1] >= 0 else -1.0, 1.0])
    if np.linalg.det(R) * d[0] * d[1] < 0:
        d[2] = -1.0
    K = K * d[None, :]
    R = d[:, None] * R
    c = -np.linalg.solve(M, P[:, 3])
    return K, R, c


def load_K_Rt_from_P(filename: Optional[str], P: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """rend_util.py:36-57: (intrinsics [4,4] float64 with K / K[2,2] in the corner, pose [4,4] float32 camera-to-world).
    With ``P=None`` the 3x4 matrix is read from a text file (an optional header line, then three rows of four numbers)."""
    if P is None:
        lines = open(filename).read().splitlines()
        if len(lines) == 4:
            lines = lines[1:]
        P = np.asarray([[float(v) for v in ln.split(" ")[:4]] for ln in lines], dtype=np.float32)
    K, R, c = decompose_projection(np.asarray(P)[:3, :4])
    intrinsics = np.eye(4)
    intrinsics[:3, :3] = K / K[2, 2]
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = R.T
    pose[:3, 3] = c
    return intrinsics, pose


def read_dtu_cameras(cam_file: str, n_images: Optional[int] = None,
                     scale_hw: Tuple[float, float] = (1.0, 1.0)) -> Dict[str, np.ndarray]:
    """``cameras.npz`` as dtu.py:78-86, 107-120 reads it.  ``scale_hw`` = (training height / image height, training
    width / image width) multiplies the intrinsics rows (dtu.py:117-118).  Returns float32 ``intrinsics [n,4,4]``,
    ``pose [n,4,4]``, ``scale_mats [n,4,4]``, ``world_mats [n,4,4]`` and ``scale_factor`` = scale_mats[0][0,0]
    (dtu.py:105)."""
    cams = np.load(cam_file)
    if n_images is None:
        n_images = sum(1 for k in cams.files if k.startswith("world_mat_") and not k.startswith("world_mat_inv_"))
    scale_mats = [cams["scale_mat_%d" % i].astype(np.float32) for i in range(n_images)]
    world_mats = [cams["world_mat_%d" % i].astype(np.float32) for i in range(n_images)]
    intr, pose = [], []
    for sm, wm in zip(scale_mats, world_mats):
        K4, c2w = load_K_Rt_from_P(None, (wm @ sm)[:3, :4])
        K4[0, :] *= scale_hw[1]
        K4[1, :] *= scale_hw[0]
        intr.append(K4.astype(np.float32))
        pose.append(c2w.astype(np.float32))
    return {"intrinsics": np.stack(intr), "pose": np.stack(pose), "scale_mats": np.stack(scale_mats),
            "world_mats": np.stack(world_mats), "scale_factor": float(scale_mats[0][0, 0])}


def camera_json_dict(poses: Sequence[np.ndarray], focals: Sequence[float], wh: Tuple[int, int],
                     files_path: Sequence[str]) -> Dict:
    """dust3r_inference_own.py:161-181 (``save_json``): principal point at the integer image centre, frames sorted by
    file path."""
    if len(poses) != len(files_path):
        raise ValueError("camera_json_dict: one pose per file")
    frames = [{"file_path": f, "transform_matrix": np.asarray(p, dtype=np.float64).tolist()} for p, f in zip(poses, files_path)]
    return {"fl_x": float(focals[0]), "fl_y": float(focals[1]), "cx": int(wh[0]) // 2, "cy": int(wh[1]) // 2,
            "w": int(wh[0]), "h": int(wh[1]), "frames": sorted(frames, key=lambda fr: fr["file_path"])}


def write_camera_json(path: str, poses, focals, wh, files_path) -> None:
    with open(path, "w") as f:
        json.dump(camera_json_dict(poses, focals, wh, files_path), f, indent=4)


def read_camera_json(path: str, keep: Optional[Iterable[str]] = None,
                     scale_hw: Tuple[float, float] = (1.0, 1.0)) -> Dict[str, object]:
    """own_data.py:55-76 / mip_nerf.py:73-150: one shared pinhole matrix (rows scaled by ``scale_hw`` as
    mip_nerf.py:107-108 does for a resized image) and the frames' camera-to-world matrices in file order; ``keep`` (file
    names, compared with the last path component as mip_nerf.py:116) selects the training views of a scene.
    Returns float32 ``intrinsics [n,4,4]`` (the reference's 3x3 in the corner of an identity: the renderer reads only
    fx, fy, cx, cy, skew), ``pose [n,4,4]``, ``img_res`` = (h, w), ``names``."""
    with open(path, "r") as f:
        data = json.load(f)
    K = np.eye(4)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = data["fl_x"], data["fl_y"], data["cx"], data["cy"]
    K[0, :] *= scale_hw[1]
    K[1, :] *= scale_hw[0]
    keep = None if keep is None else set(keep)
    poses: List[np.ndarray] = []
    names: List[str] = []
    for frame in data["frames"]:
        name = frame["file_path"].split("/")[-1]
        if keep is not None and name not in keep:
            continue
        c2w = np.asarray(frame["transform_matrix"], dtype=np.float64)
        if c2w.shape == (3, 4):
            c2w = np.vstack([c2w, [0.0, 0.0, 0.0, 1.0]])
        if c2w.shape != (4, 4):
            raise ValueError(f"read_camera_json: frame '{name}' has a {c2w.shape} transform_matrix")
        poses.append(c2w.astype(np.float32))
        names.append(name)
    n = len(poses)
    return {"intrinsics": np.repeat(K[None].astype(np.float32), n, axis=0),
            "pose": np.stack(poses) if n else np.zeros((0, 4, 4), np.float32),
            "img_res": (int(data["h"]), int(data["w"])), "names": names}
