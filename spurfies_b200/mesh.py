"""the mesh-extraction caller of the hot path -- SDF volume on a regular grid, kept on the device.

Mirrors ``spurfies/utils/plots.py``: ``get_grid_uniform`` (:289-300), ``get_grid`` (:302-333) and the SDF-evaluation
half of ``get_surface_by_grid`` (:188-287; ``higher_res=False`` is the only mode ``eval_spurfies.py:151-176`` uses, the
PCA-aligned second pass of ``higher_res=True`` takes the low-resolution mesh's surface samples from the caller).
The reference materialises all grid points, evaluates ``model.get_sdf_eval`` (pointneus_disent.py:249-298) in chunks
of 100 000 and copies every chunk to the host; here ``sdf_volume`` generates the points on the fly, skips everything
outside the dilated occupancy of the neural points (those values are the constant 1000) and leaves the volume in HBM.
Grid slabs shard across ranks with no collective (``rank`` / ``world``).

Marching cubes itself (``skimage.measure.marching_cubes`` on the host in the reference) is not part of this path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream
from .dist import shard_range
from .fields import SlotSet, geo_sdf_raw, set_precision

NO_NEIGHBOUR = 1000.0   # pointneus_disent.py:296


def get_grid_uniform(resolution: int, grid_boundary=(-2.0, 2.0)) -> Dict:
    """plots.py:289-300 without materialising ``grid_points``."""
    x = np.linspace(grid_boundary[0], grid_boundary[1], resolution)
    return {"shortest_axis_length": 2.0, "xyz": [x, x, x], "shortest_axis_index": 0}


def get_grid(points, resolution: int, input_min=None, input_max=None, eps: float = 0.1) -> Dict:
    """plots.py:302-333 without materialising ``grid_points``: the shortest bounding-box axis gets `resolution`
    samples, the other two the same spacing."""
    if input_min is None or input_max is None:
        input_min = torch.min(points, dim=0)[0].squeeze().cpu().numpy()
        input_max = torch.max(points, dim=0)[0].squeeze().cpu().numpy()
    input_min, input_max = np.asarray(input_min), np.asarray(input_max)   # keep the caller's dtype, as the reference does
    s = int(np.argmin(input_max - input_min))
    axes = [None, None, None]
    axes[s] = np.linspace(input_min[s] - eps, input_max[s] + eps, resolution)
    length = np.max(axes[s]) - np.min(axes[s])
    step = length / (axes[s].shape[0] - 1)
    for a in range(3):
        if a != s:
            axes[a] = np.arange(input_min[a] - eps, input_max[a] + step + eps, step)
    return {"shortest_axis_length": length, "xyz": axes, "shortest_axis_index": s}


def grid_points(xyz: Sequence[np.ndarray]) -> torch.Tensor:
    """The reference's materialised point list (plots.py:328-329), for tests and small grids only."""
    xx, yy, zz = np.meshgrid(*xyz)
    return torch.tensor(np.vstack([xx.ravel(), yy.ravel(), zz.ravel()]).T, dtype=torch.float)


def cyclic_local_count(G: int, rank: int, world: int, block: int) -> int:
    """Number of grid points of `rank` under the block-cyclic distribution (block b of `block` points -> rank b % world)."""
    nb = (G + block - 1) // block
    mine = (nb - rank + world - 1) // world if rank < nb else 0
    n = mine * block
    if mine > 0 and (nb - 1) % world == rank:      # the (possibly partial) last block is ours
        n -= nb * block - G
    return n


def cyclic_global_index(local: torch.Tensor, rank: int, world: int, block: int) -> torch.Tensor:
    """Global grid index of a rank's local indices (spf_grid_points_mask_cyclic)."""
    local = local.to(torch.int64)
    return ((local // block) * world + rank) * block + local % block


@torch.no_grad()
def sdf_volume(model, xyz: Sequence[np.ndarray], chunk: int = 1 << 24, rank: int = 0, world: int = 1,
               out: Optional[torch.Tensor] = None, cyclic_block: Optional[int] = None,
               affine: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, Tuple[int, int]]:
    """SDF of ``model.get_sdf_eval`` at the grid points of axes ``xyz`` in the reference's order (index =
    (iy * nx + ix) * nz + iz).  Returns (flat fp32 device tensor over this rank's contiguous index range, (lo, hi)).
    Reshape the full volume as ``[ny, nx, nz]`` (plots.py:259-261 then transposes it to [nx, ny, nz]).

    ``cyclic_block`` (multi-GPU): instead of a contiguous slab the rank owns every world-th block of that many grid points
    (``cyclic_global_index`` maps its consecutive local indices to grid indices) -- a contiguous slab that crosses the
    object costs several times one that does not; the returned range is then (0, number of local points).

    ``affine = (M [3,3], c [3])``: the query points are ``M @ (x, y, z) + c`` (the PCA-aligned grid of
    ``get_surface_by_grid(higher_res=True)``, plots.py:240-246, generated on the fly like the axis-aligned one)."""
    set_precision(model.precision)
    dev = model.neural_pts.device
    ax = [torch.as_tensor(np.asarray(a, dtype=np.float64)).to(torch.float32).to(dev).contiguous() for a in xyz]
    nx, ny, nz = (int(a.numel()) for a in ax)
    G = nx * ny * nz
    if cyclic_block is None or world == 1:
        lo, hi = shard_range(G, rank, world)
        cyc = (1, 1, 0)
    else:
        lo, hi = 0, cyclic_local_count(G, rank, world, int(cyclic_block))
        cyc = (int(cyclic_block), int(world), int(rank))
    if out is None:
        out = torch.empty(hi - lo, dtype=torch.float32, device=dev)
    assert out.numel() == hi - lo and out.is_cuda and out.dtype == torch.float32
    grid = model._grid()
    pack = model._pack()
    chunk = int(min(chunk, max(hi - lo, 1)))
    idx = torch.empty(chunk, dtype=torch.int32, device=dev)
    pts = torch.empty(chunk, 3, dtype=torch.float32, device=dev)
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    r2 = grid.radius2(model.conf.r)
    aff = None
    if affine is not None:
        aff = torch.cat([affine[0].reshape(9), affine[1].reshape(3)]).to(device=dev, dtype=torch.float32).contiguous()
    for c0 in range(lo, hi, chunk):
        n = min(chunk, hi - c0)
        vol = out[c0 - lo:c0 - lo + n]
        call("spf_grid_points_mask_cyclic", C.byref(grid.handle), ptr(ax[0]), ptr(ax[1]), ptr(ax[2]), nx, ny, nz, c0, n,
             cyc[0], cyc[1], cyc[2], ptr(aff), NO_NEIGHBOUR, ptr(vol), ptr(idx), ptr(pts), ptr(counter), chunk, stream())
        m = int(counter.item())            # one 4-byte readback per chunk (the reference copies the whole chunk)
        if m == 0:
            continue
        q = pts[:m]
        pidx = torch.empty(m, model.conf.k, dtype=torch.int32, device=dev)
        call("spf_knn_points", C.byref(grid.handle), ptr(q), m, model.conf.k, r2, ptr(pidx), stream())
        slots = SlotSet(pidx, "mesh")
        sdf, _, _ = geo_sdf_raw(pack, slots, q, model.neural_pts, model.neural_feats_geometry.detach(), model.conf.rbf,
                                False, False, fill=NO_NEIGHBOUR)
        call("spf_scatter_f32", ptr(idx), ptr(sdf), m, ptr(vol), stream())
    return out, (lo, hi)


def pca_alignment(recon_pc: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """plots.py:222-231: principal axes `vecs` (rows) and mean of the point cloud sampled from the low-resolution mesh."""
    s_mean = recon_pc.mean(dim=0)
    d = recon_pc - s_mean
    s_cov = torch.mm(d.transpose(0, 1), d)
    # (3x3: decomposed on the host -- LAPACK's ordering / signs, not a device library's that may differ between versions)
    vecs = torch.view_as_real(torch.linalg.eig(s_cov.cpu())[1].transpose(0, 1))[:, :, 0].to(recon_pc.device)
    if torch.det(vecs) < 0:
        vecs = torch.mm(torch.tensor([[1, 0, 0], [0, 0, 1], [0, 1, 0]], device=vecs.device, dtype=vecs.dtype), vecs)
    return vecs, s_mean


def get_surface_by_grid(grid_params, model, resolution: int = 100, level: float = 0.0, higher_res: bool = False,
                        chunk: int = 1 << 24, recon_pc: Optional[torch.Tensor] = None) -> Dict:
    """plots.py:188-287 up to the marching-cubes call: returns the device SDF volume in the layout the reference hands to
    ``measure.marching_cubes`` ([nx, ny, nz]), the grid spacing and origin, and whether the level set crosses it.

    ``higher_res=True`` (plots.py:196-246): the second, PCA-aligned pass.  Its first half -- marching cubes of a 100^3
    volume (``get_surface_by_grid(..., resolution=100)`` gives that volume), largest component, 10 000 surface samples --
    is skimage / trimesh work outside the path; hand its samples in as ``recon_pc`` [P,3].  This function then does the
    alignment (``pca_alignment``), builds the aligned grid and queries the SDF at the ROTATED grid points
    ``vecs^T p + mean`` (generated on the fly, ``sdf_volume(affine=)``); the result also carries ``vecs`` / ``mean`` and
    the first grid point that plots.py:270-275 uses to map the vertices back."""
    if higher_res:
        if recon_pc is None:
            raise ValueError("higher_res=True needs recon_pc: the points sampled from the low-resolution mesh "
                             "(plots.py:198-220; marching cubes and mesh sampling are outside this package)")
        dev = model.neural_pts.device
        vecs, s_mean = pca_alignment(recon_pc.to(dev).float())
        helper = torch.mm(recon_pc.to(dev).float() - s_mean, vecs.transpose(0, 1))          # plots.py:231-232
        grid = get_grid(helper.cpu(), resolution, eps=0.01)
        x, y, z = grid["xyz"]
        M = vecs.transpose(0, 1).contiguous()                                                 # plots.py:243-245
        flat, _ = sdf_volume(model, grid["xyz"], chunk=chunk, affine=(M, s_mean))
        first = torch.mv(M, torch.tensor([x[0], y[0], z[0]], dtype=torch.float32, device=dev)) + s_mean
    else:
        gp = np.asarray(grid_params, dtype=np.float64) * np.array([[1.5], [1.0]])   # plots.py:189
        grid = get_grid(None, resolution, input_min=gp[0], input_max=gp[1], eps=0.0)
        x, y, z = grid["xyz"]
        flat, _ = sdf_volume(model, grid["xyz"], chunk=chunk)
    vol = flat.view(len(y), len(x), len(z)).permute(1, 0, 2)
    lo, hi = torch.aminmax(vol)
    spacing = float(x[2] - x[1])
    out = {"volume": vol, "spacing": (spacing, spacing, spacing), "origin": (float(x[0]), float(y[0]), float(z[0])),
           "has_surface": not (float(lo) > level or float(hi) < level), "xyz": grid["xyz"]}
    if higher_res:
        out.update(vecs=vecs, mean=s_mean, first_grid_point=first)
    return out
