// mesh.cu -- SURVEY 8(f2): the SDF-grid query that feeds marching cubes (spurfies/utils/plots.py:188-287,
// eval_spurfies.py:140-194; BASELINE config "Marching-cubes SDF grid query 512^3 via kNN + prior MLP").
//
// The reference materialises every grid point ([G,3] fp32 on the device), pushes 100 000-point chunks through
// get_sdf_eval and copies each chunk's result to the host.  Here the grid points are never materialised: one pass
// generates them from the three axis vectors, writes the "no neighbour" constant (1000, pointneus_disent.py:296) to
// the volume and compacts the few points that fall inside the dilated occupancy of the neural points (knnquery.cu:
// 171-196); only those go through spf_knn_points + spf_sdf_fwd_*, and spf_scatter_f32 writes their SDF back.
#include "common.cuh"

// reference point order (plots.py:328-329): np.meshgrid(x, y, z) with the default 'xy' indexing, raveled:
//   linear index = (iy * nx + ix) * nz + iz,  point = (x[ix], y[iy], z[iz])
__global__ void k_grid_points_mask(GridDev g, const float* __restrict__ xs, const float* __restrict__ ys,
                                   const float* __restrict__ zs, int nx, int ny, int nz, long long lo, long long count,
                                   long long cyc_block, int cyc_world, int cyc_rank, const float* __restrict__ affine,
                                   float fill, float* __restrict__ vol, int* __restrict__ idx_out,
                                   float* __restrict__ pts_out, int* __restrict__ counter, int cap) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool h = false;
  float x = 0.f, y = 0.f, z = 0.f;
  if (t < count) {
    // block-cyclic distribution over ranks (cyc_world == 1: identity): local index j = lo + t lives in this rank's
    // (j / block)-th block, which is global block (j / block) * world + rank
    const long long j = lo + t;
    const long long i = cyc_world == 1 ? j : ((j / cyc_block) * cyc_world + cyc_rank) * cyc_block + (j % cyc_block);
    const int iz = (int)(i % nz);
    const long long r = i / nz;
    const int ix = (int)(r % nx), iy = (int)(r / nx);
    x = xs[ix]; y = ys[iy]; z = zs[iz];
    if (affine) {   // PCA-aligned grid of the higher_res pass (plots.py:240-246): p = M (x, y, z) + c, M row-major
      const float gx = x, gy = y, gz = z;
      x = affine[0] * gx + affine[1] * gy + affine[2] * gz + affine[9];
      y = affine[3] * gx + affine[4] * gy + affine[5] * gz + affine[10];
      z = affine[6] * gx + affine[7] * gy + affine[8] * gz + affine[11];
    }
    vol[t] = fill;
    int cx, cy, cz;
    const int v = voxel_of(g, x, y, z, cx, cy, cz);
    h = v >= 0 && g.hit[v];
  }
  const unsigned m = __ballot_sync(SPF_FULL, h);
  if (m == 0) return;
  int base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(SPF_FULL, base, __ffs(m) - 1);
  if (h) {
    const int o = base + __popc(m & ((1u << lane) - 1));
    if (o < cap) {
      idx_out[o] = (int)t;
      pts_out[3 * (size_t)o] = x; pts_out[3 * (size_t)o + 1] = y; pts_out[3 * (size_t)o + 2] = z;
    }
  }
}

__global__ void k_scatter_f32(const int* __restrict__ idx, const float* __restrict__ vals, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[idx[i]] = vals[i];
}

extern "C" int spf_grid_points_mask_cyclic(const spf_grid* g, const float* xs, const float* ys, const float* zs, int32_t nx,
                                           int32_t ny, int32_t nz, int64_t lo, int64_t count, int64_t block, int32_t world,
                                           int32_t rank, const float* affine, float fill, float* vol, int32_t* idx_out,
                                           float* pts_out, int32_t* counter, int32_t cap, void* stream_) {
  if (!g || !xs || !ys || !zs || !vol || !idx_out || !pts_out || !counter) return SPF_ERR_INVALID;
  if (nx <= 0 || ny <= 0 || nz <= 0 || lo < 0 || count < 0 || block < 1 || world < 1 || rank < 0 || rank >= world)
    return SPF_ERR_INVALID;
  if (count > 0x7fffffffLL) return SPF_ERR_UNSUPPORTED;   // chunk-local indices are int32
  if (count > 0) {   // the last local index must map inside the grid
    const int64_t j = lo + count - 1;
    const int64_t i = world == 1 ? j : ((j / block) * world + rank) * block + (j % block);
    if (i >= (int64_t)nx * ny * nz) return SPF_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream_;
  SPF_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st), "grid_points_mask memset");
  if (count == 0) return SPF_OK;
  k_grid_points_mask<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(to_dev(g), xs, ys, zs, nx, ny, nz, lo, count, block, world,
                                                                     rank, affine, fill, vol, idx_out, pts_out, counter, cap);
  SPF_CHECK_LAUNCH("k_grid_points_mask");
  return SPF_OK;
}

extern "C" int spf_grid_points_mask(const spf_grid* g, const float* xs, const float* ys, const float* zs, int32_t nx,
                                    int32_t ny, int32_t nz, int64_t lo, int64_t count, float fill, float* vol,
                                    int32_t* idx_out, float* pts_out, int32_t* counter, int32_t cap, void* stream_) {
  return spf_grid_points_mask_cyclic(g, xs, ys, zs, nx, ny, nz, lo, count, 1, 1, 0, nullptr, fill, vol, idx_out, pts_out, counter,
                                     cap, stream_);
}

extern "C" int spf_scatter_f32(const int32_t* idx, const float* vals, int32_t n, float* out, void* stream_) {
  if (n <= 0) return SPF_OK;
  if (!idx || !vals || !out) return SPF_ERR_INVALID;
  k_scatter_f32<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(idx, vals, n, out);
  SPF_CHECK_LAUNCH("k_scatter_f32");
  return SPF_OK;
}
