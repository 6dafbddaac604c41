"""neural-point ingestion -- the step before the hot path.

Mirrors ``spurfies/model/utils.py:6-88``: ``construct_vox_points_closest`` (voxel-downsample to the input point closest
to each occupied voxel's centroid), ``voxelize`` and ``load_neural_points`` (PLY with ``x, y, z`` and optional
``red, green, blue`` vertex properties, written by ``dust3r_inference_own.py:73-86``).  The reference needs
``plyfile`` and ``torch_scatter``; here the PLY codec is numpy and the voxel pass is ``spf_voxelize_closest``
(deterministic: fixed-point centroid sums, ties broken by the smallest point index).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
              "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
              "double": "f8", "float64": "f8"}


def read_ply(path: str) -> Dict[str, np.ndarray]:
    """Vertex properties of a PLY file (ascii, binary_little_endian or binary_big_endian) as {name: array}."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements, cur = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                cur = {"name": tok[1], "count": int(tok[2]), "props": []}
                elements.append(cur)
            elif tok[0] == "property":
                if tok[1] == "list":
                    cur["props"].append(("list", tok[2], tok[3], tok[4]))
                else:
                    cur["props"].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if not elements or elements[0]["name"] != "vertex":
            raise ValueError(f"{path}: the first element must be `vertex`")
        v = elements[0]
        if any(p[0] == "list" for p in v["props"]):
            raise ValueError(f"{path}: list properties on vertices are not supported")
        if fmt == "ascii":
            rows = np.loadtxt(f, max_rows=v["count"], ndmin=2)
            return {name: rows[:, i].astype(t) for i, (name, t) in enumerate(v["props"])}
        end = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(name, end + t) for name, t in v["props"]])
        rec = np.frombuffer(f.read(dt.itemsize * v["count"]), dtype=dt, count=v["count"])
        return {name: np.ascontiguousarray(rec[name]) for name, _ in v["props"]}


def write_ply(path: str, pts: np.ndarray, colors: Optional[np.ndarray] = None) -> None:
    """Binary little-endian PLY: float x, y, z (+ uchar red, green, blue), the layout dust3r_inference_own.py:73-86 writes."""
    pts = np.asarray(pts, dtype=np.float32)
    fields = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")]
    if colors is not None:
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
    rec = np.empty(len(pts), dtype=np.dtype(fields))
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    if colors is not None:
        c = np.asarray(colors).astype(np.uint8)
        rec["red"], rec["green"], rec["blue"] = c[:, 0], c[:, 1], c[:, 2]
    names = {"<f4": "float", "u1": "uchar"}
    head = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % len(pts)
    head += "".join("property %s %s\n" % (names[t], n) for n, t in fields) + "end_header\n"
    with open(path, "wb") as f:
        f.write(head.encode("ascii"))
        f.write(rec.tobytes())


def construct_vox_points_closest(xyz_val: torch.Tensor, vox_res, partition_xyz=None, space_min=None, space_max=None
                                 ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """utils.py:6-37 -> (xyz_centroid [V,3], sparse_grid_idx i32 [V,3], min_idx i64 [V]); voxels in sorted (x, y, z)
    order like torch.unique(dim=0).  Only the default call form (no partition, bounds from the points) is used by the
    reference (utils.py:49-55) and supported here."""
    if partition_xyz is not None or space_min is not None or space_max is not None:
        raise NotImplementedError("construct_vox_points_closest: only the bounds-from-points form is used by voxelize")
    if not xyz_val.is_cuda:
        raise _lib.SpfError("construct_vox_points_closest needs a CUDA tensor (there is no CPU fallback)")
    xyz = xyz_val.float().contiguous()
    n = xyz.shape[0]
    # utils.py:11-15, 22: the same fp32 torch ops, evaluated once
    xyz_min, xyz_max = torch.min(xyz, dim=-2)[0], torch.max(xyz, dim=-2)[0]
    space_edge = torch.max(xyz_max - xyz_min) * 1.05
    xyz_mid = (xyz_max + xyz_min) / 2
    smin = xyz_mid - space_edge / 2
    vox_sz = space_edge / vox_res
    host = torch.cat([smin, vox_sz.reshape(1)]).cpu()      # one 16-byte readback per cloud
    M = int(vox_res) + 1
    dev = xyz.device
    ws = torch.empty(_lib.lib.spf_voxelize_workspace_bytes(M), dtype=torch.uint8, device=dev)
    cap = max(n, 1)
    min_idx = torch.empty(cap, dtype=torch.int64, device=dev)
    centroid = torch.empty(cap, 3, dtype=torch.float32, device=dev)
    gidx = torch.empty(cap, 3, dtype=torch.int32, device=dev)
    n_out = torch.zeros(2, dtype=torch.int32, device=dev)
    call("spf_voxelize_closest", ptr(xyz), n, float(host[0]), float(host[1]), float(host[2]), float(host[3]), M,
         ptr(min_idx), ptr(centroid), ptr(gidx), cap, ptr(n_out), ptr(ws), ws.numel(), stream())
    nv, bad = (int(v) for v in n_out.cpu())
    if bad:
        raise _lib.SpfError(f"voxelize: {bad} points fell outside the {M}^3 voxel table")
    return centroid[:nv], gidx[:nv], min_idx[:nv]


def voxelize(pointcloud: Union[torch.Tensor, List[torch.Tensor]], vox_res) -> Tuple[torch.Tensor, torch.Tensor]:
    """utils.py:39-59 -> (down-sampled points, indices of the kept points in the last cloud)."""
    clouds = [pointcloud] if not isinstance(pointcloud, list) else pointcloud
    holder = torch.zeros([0, 3], dtype=clouds[0].dtype, device="cuda")
    sampled = None
    for i, pts in enumerate(clouds):
        vox_res = vox_res // (1.5 ** i)
        src = pts.cuda() if len(pts) < 80000000 else pts[:: (len(pts) // 80000000 + 1), ...].cuda()
        _, _, sampled = construct_vox_points_closest(src, vox_res)
        holder = torch.cat([holder, pts.cuda()[sampled, :]], dim=0)
    return holder, sampled


def load_neural_points(path: str, vox_res=None) -> Dict[str, torch.Tensor]:
    """utils.py:61-88: {"pts": [N,3] (cuda), "colors": [N,3] (cuda, only if the file has red/green/blue)}."""
    ply = read_ply(path)
    pointcloud = torch.from_numpy(np.stack([ply["x"], ply["y"], ply["z"]], axis=-1)).to("cuda")
    idx = None
    if vox_res is not None:
        pointcloud, idx = voxelize(pointcloud, vox_res)
    out = {"pts": pointcloud}
    if "red" in ply:
        color = torch.from_numpy(np.stack([ply["red"], ply["green"], ply["blue"]], axis=-1)).to("cuda")
        out["colors"] = color[idx, :] if idx is not None else color
    return out
