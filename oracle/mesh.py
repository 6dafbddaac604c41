"""TEST INFRASTRUCTURE ONLY (never imported by spurfies_b200/): CPU restatement of the grid construction of the
reference's mesh extraction, spurfies/utils/plots.py:289-333 (`get_grid_uniform`, `get_grid`) and of the SDF-volume
half of `get_surface_by_grid` (:188-287, higher_res=False), with the points MATERIALISED exactly as the reference does
(np.meshgrid -> vstack -> float32 tensor) and pushed through the oracle's point SDF in `splitn` chunks.

plots.py imports skimage / trimesh (absent here), so it cannot be imported to generate fixtures: parity of this
restatement is pinned only through oracle/hotpath.py::point_sdf, which IS pinned by the reference-generated golden
vectors (tests/golden/hotpath_dtu4k.pt: `sdf_importance`)."""
import numpy as np
import torch

from . import hotpath as H


def get_grid_uniform(resolution, grid_boundary=(-2.0, 2.0)):  # plots.py:289-300
    x = np.linspace(grid_boundary[0], grid_boundary[1], resolution)
    y = x
    z = x
    xx, yy, zz = np.meshgrid(x, y, z)
    pts = torch.tensor(np.vstack([xx.ravel(), yy.ravel(), zz.ravel()]).T, dtype=torch.float)
    return {"grid_points": pts, "shortest_axis_length": 2.0, "xyz": [x, y, z], "shortest_axis_index": 0}


def get_grid(points, resolution, input_min=None, input_max=None, eps=0.1):  # plots.py:302-333
    if input_min is None or input_max is None:
        input_min = torch.min(points, dim=0)[0].squeeze().numpy()
        input_max = torch.max(points, dim=0)[0].squeeze().numpy()
    bounding_box = input_max - input_min
    shortest_axis = np.argmin(bounding_box)
    if shortest_axis == 0:
        x = np.linspace(input_min[shortest_axis] - eps, input_max[shortest_axis] + eps, resolution)
        length = np.max(x) - np.min(x)
        y = np.arange(input_min[1] - eps, input_max[1] + length / (x.shape[0] - 1) + eps, length / (x.shape[0] - 1))
        z = np.arange(input_min[2] - eps, input_max[2] + length / (x.shape[0] - 1) + eps, length / (x.shape[0] - 1))
    elif shortest_axis == 1:
        y = np.linspace(input_min[shortest_axis] - eps, input_max[shortest_axis] + eps, resolution)
        length = np.max(y) - np.min(y)
        x = np.arange(input_min[0] - eps, input_max[0] + length / (y.shape[0] - 1) + eps, length / (y.shape[0] - 1))
        z = np.arange(input_min[2] - eps, input_max[2] + length / (y.shape[0] - 1) + eps, length / (y.shape[0] - 1))
    else:
        z = np.linspace(input_min[shortest_axis] - eps, input_max[shortest_axis] + eps, resolution)
        length = np.max(z) - np.min(z)
        x = np.arange(input_min[0] - eps, input_max[0] + length / (z.shape[0] - 1) + eps, length / (z.shape[0] - 1))
        y = np.arange(input_min[1] - eps, input_max[1] + length / (z.shape[0] - 1) + eps, length / (z.shape[0] - 1))
    xx, yy, zz = np.meshgrid(x, y, z)
    pts = torch.tensor(np.vstack([xx.ravel(), yy.ravel(), zz.ravel()]).T, dtype=torch.float)
    return {"grid_points": pts, "shortest_axis_length": length, "xyz": [x, y, z], "shortest_axis_index": shortest_axis}


def surface_volume(p, grid, grid_params, resolution=100, splitn=100000):
    """plots.py:188-190, 236, 250-261: SDF at every grid point -> volume [nx, ny, nz] as handed to marching_cubes."""
    gp = np.asarray(grid_params, dtype=np.float64) * [[1.5], [1.0]]
    g = get_grid(None, resolution, input_min=gp[0], input_max=gp[1], eps=0.0)
    z = []
    with torch.no_grad():
        for pnts in torch.split(g["grid_points"], splitn, dim=0):
            z.append(H.point_sdf(p, grid, pnts).numpy())
    z = np.concatenate(z, axis=0).astype(np.float32)
    x, y, zz = g["xyz"]
    return z.reshape(y.shape[0], x.shape[0], zz.shape[0]).transpose([1, 0, 2]), g


def aligned_surface_volume(p, grid, recon_pc, resolution=100, splitn=100000):
    """plots.py:222-261 with higher_res=True, from the surface samples `recon_pc` of the low-resolution mesh on: PCA
    alignment, aligned grid (materialised), rotation of the grid points back into the scene frame, SDF of every point ->
    (volume [nx, ny, nz], grid dict, vecs, s_mean, rotated grid points)."""
    recon_pc = recon_pc.float()
    s_mean = recon_pc.mean(dim=0)
    s_cov = recon_pc - s_mean
    s_cov = torch.mm(s_cov.transpose(0, 1), s_cov)
    vecs = torch.view_as_real(torch.linalg.eig(s_cov)[1].transpose(0, 1))[:, :, 0]
    if torch.det(vecs) < 0:
        vecs = torch.mm(torch.tensor([[1, 0, 0], [0, 0, 1], [0, 1, 0]]).float(), vecs)
    helper = torch.bmm(vecs.unsqueeze(0).repeat(recon_pc.shape[0], 1, 1), (recon_pc - s_mean).unsqueeze(-1)).squeeze()
    g = get_grid(helper, resolution, eps=0.01)
    pts = []
    for pnts in torch.split(g["grid_points"], splitn, dim=0):
        pts.append(torch.bmm(vecs.unsqueeze(0).repeat(pnts.shape[0], 1, 1).transpose(1, 2), pnts.unsqueeze(-1)).squeeze() + s_mean)
    pts = torch.cat(pts, dim=0)
    z = []
    with torch.no_grad():
        for pnts in torch.split(pts, splitn, dim=0):
            z.append(H.point_sdf(p, grid, pnts).numpy())
    z = np.concatenate(z, axis=0).astype(np.float32)
    x, y, zz = g["xyz"]
    return z.reshape(y.shape[0], x.shape[0], zz.shape[0]).transpose([1, 0, 2]), g, vecs, s_mean, pts
