"""TEST INFRASTRUCTURE ONLY (never imported by spurfies_b200/): CPU restatement of the reference's neural-point
voxel down-sampling, spurfies/model/utils.py:6-59 (`construct_vox_points_closest`, `voxelize`).

The reference uses torch_scatter's scatter_mean / scatter_min (absent in this image; on CUDA they run on float atomics,
so the centroid's last bits and the tie-breaks of scatter_min are not reproducible run to run even upstream).
PINNED: tests/golden/ingest.pt holds the outputs of the reference's own utils.py functions, imported from
/root/reference with torch_scatter replaced by a pure-torch shim of its documented CPU semantics
(tests/golden/make_golden_ingest.py); tests/test_oracle_ingest.py checks this restatement against it (voxel set and
order exact, kept point per voxel identical, centroid within one fp32 ulp of the fp32-summed reference), and
tests/test_gpu_ingest.py checks the kernels against the same fixture.  scatter_mean -> index_add / count (here in
fp64, rounded once), scatter_min -> amin reduce + first index attaining it; everything else is the same torch ops in
the same order."""
import torch


def construct_vox_points_closest(xyz_val, vox_res):  # utils.py:6-37 (bounds-from-points form)
    xyz = xyz_val
    xyz_min, xyz_max = torch.min(xyz, dim=-2)[0], torch.max(xyz, dim=-2)[0]
    space_edge = torch.max(xyz_max - xyz_min) * 1.05
    xyz_mid = (xyz_max + xyz_min) / 2
    space_min = xyz_mid - space_edge / 2
    # The reference evaluates this on CUDA tensors (utils.py:49-55 moves the cloud to the GPU first), where torch divides a
    # tensor by a Python scalar as `a * (1 / b)` in fp32 (ATen div_true_kernel_cuda), not as an IEEE division: restate that.
    construct_vox_sz = space_edge * (torch.tensor(1.0, dtype=torch.float32) / torch.tensor(float(vox_res), dtype=torch.float32))
    xyz_shift = xyz - space_min[None, ...]
    sparse_grid_idx, inv_idx = torch.unique(torch.floor(xyz_shift / construct_vox_sz[None, ...]).to(torch.int32), dim=0,
                                            return_inverse=True)
    V = sparse_grid_idx.shape[0]
    cnt = torch.zeros(V, dtype=torch.float64).index_add_(0, inv_idx, torch.ones(len(xyz), dtype=torch.float64))
    xyz_centroid = (torch.zeros(V, 3, dtype=torch.float64).index_add_(0, inv_idx, xyz_val.double()) / cnt[:, None]).float()
    xyz_centroid_prop = xyz_centroid[inv_idx, :]
    xyz_residual = torch.norm(xyz_val - xyz_centroid_prop, dim=-1)
    best = torch.full((V,), float("inf")).scatter_reduce_(0, inv_idx, xyz_residual, reduce="amin")
    is_min = xyz_residual == best[inv_idx]
    n = len(xyz)
    cand = torch.where(is_min, torch.arange(n), torch.full((n,), n))
    min_idx = torch.full((V,), n, dtype=torch.long).scatter_reduce_(0, inv_idx, cand, reduce="amin")
    return xyz_centroid, sparse_grid_idx, min_idx, xyz_residual, inv_idx


def voxelize(pointcloud, vox_res):  # utils.py:39-59, single cloud
    _, _, idx, _, _ = construct_vox_points_closest(pointcloud, vox_res)
    return pointcloud[idx, :], idx
