// ingest.cu -- SURVEY 8(f3): neural-point ingestion, the step before the hot path.
// construct_vox_points_closest (spurfies/model/utils.py:6-37): voxel-downsample a point cloud to ONE point per occupied
// voxel -- the input point closest to the voxel's centroid -- with voxels enumerated in sorted (x, y, z) order
// (torch.unique(dim=0)).  The reference does unique + torch_scatter scatter_mean / scatter_min (float atomics, so its
// centroid bits and its tie-breaks vary run to run); here:
//   k_vox_accum   : per point, voxel index with the reference's fp32 arithmetic; count and a 2^-32 fixed-point int64
//                   coordinate sum per voxel (integer atomics: order-independent, deterministic)
//   k_vox_select  : per point, residual |p - centroid| (fp32, as torch.norm) -> 64-bit atomicMin of (residual bits, index)
//                   (deterministic: smallest residual, then smallest point index)
//   k_vox_count / k_vox_scan / k_vox_emit : ordered compaction of the occupied voxels
// over a dense voxel table (vox_res = 300 -> 27 M cells).
#include "common.cuh"

struct VoxDev {
  float sx, sy, sz;   // space_min
  float vs;           // construct_vox_sz
  int M;              // cells per axis
};

__device__ __forceinline__ long long vox_cell(const VoxDev& v, float x, float y, float z, int& ix, int& iy, int& iz) {
  // torch.floor((xyz - space_min) / construct_vox_sz).to(int32)   (utils.py:24-26)
  ix = (int)floorf(__fdiv_rn(__fsub_rn(x, v.sx), v.vs));
  iy = (int)floorf(__fdiv_rn(__fsub_rn(y, v.sy), v.vs));
  iz = (int)floorf(__fdiv_rn(__fsub_rn(z, v.sz), v.vs));
  if (ix < 0 || iy < 0 || iz < 0 || ix >= v.M || iy >= v.M || iz >= v.M) return -1;
  return ((long long)ix * v.M + iy) * v.M + iz;
}

#define FIX_SCALE 4294967296.0

__global__ void k_vox_accum(VoxDev v, const float* __restrict__ pts, int n, int* __restrict__ count,
                            long long* __restrict__ sums, int* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
  int ix, iy, iz;
  const long long c = vox_cell(v, x, y, z, ix, iy, iz);
  if (c < 0) { atomicAdd(err, 1); return; }
  atomicAdd(count + c, 1);
  atomicAdd(reinterpret_cast<unsigned long long*>(sums + 3 * c), (unsigned long long)__double2ll_rn((double)x * FIX_SCALE));
  atomicAdd(reinterpret_cast<unsigned long long*>(sums + 3 * c + 1), (unsigned long long)__double2ll_rn((double)y * FIX_SCALE));
  atomicAdd(reinterpret_cast<unsigned long long*>(sums + 3 * c + 2), (unsigned long long)__double2ll_rn((double)z * FIX_SCALE));
}

__device__ __forceinline__ float vox_centroid(const long long* sums, long long c, int a, int cnt) {
  return (float)((double)sums[3 * c + a] / ((double)cnt * FIX_SCALE));
}

__global__ void k_vox_select(VoxDev v, const float* __restrict__ pts, int n, const int* __restrict__ count,
                             const long long* __restrict__ sums, unsigned long long* __restrict__ best) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
  int ix, iy, iz;
  const long long c = vox_cell(v, x, y, z, ix, iy, iz);
  if (c < 0) return;
  const int cnt = count[c];
  const float dx = __fsub_rn(x, vox_centroid(sums, c, 0, cnt)), dy = __fsub_rn(y, vox_centroid(sums, c, 1, cnt)),
              dz = __fsub_rn(z, vox_centroid(sums, c, 2, cnt));
  const float res = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  atomicMin(best + c, ((unsigned long long)__float_as_uint(res) << 32) | (unsigned)i);
}

#define VOX_TILE 1024
__global__ void k_vox_count(const int* __restrict__ count, long long G, int* __restrict__ block_sums) {
  const long long base = (long long)blockIdx.x * VOX_TILE;
  int c = 0;
  for (int j = threadIdx.x; j < VOX_TILE; j += blockDim.x) {
    const long long i = base + j;
    c += (i < G && count[i] > 0) ? 1 : 0;
  }
  c = (int)warp_sum((float)c);   // <= 1024: exact in fp32
  __shared__ int s_w[8];
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_w[w];
    block_sums[blockIdx.x] = t;
  }
}

// exclusive scan of block_sums in place (single block, sequential over chunks of 1024), total -> n_out[0]
__global__ void k_vox_scan(int* __restrict__ block_sums, int nb, int* __restrict__ n_out) {
  __shared__ int s[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? block_sums[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nb) block_sums[i] = carry + s[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += s[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) n_out[0] = carry;
}

__global__ void k_vox_emit(VoxDev v, const int* __restrict__ count, const long long* __restrict__ sums,
                           const unsigned long long* __restrict__ best, long long G, const int* __restrict__ block_offs,
                           int cap, long long* __restrict__ min_idx, float* __restrict__ centroid,
                           int* __restrict__ grid_idx) {
  // one warp-ordered pass per 1024-cell tile: 256 threads x 4 consecutive cells
  const long long base = (long long)blockIdx.x * VOX_TILE + (long long)threadIdx.x * 4;
  int occ[4], my = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) { occ[j] = (base + j < G && count[base + j] > 0) ? 1 : 0; my += occ[j]; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int incl = warp_scan_incl_i(my, lane);
  __shared__ int s_w[8];
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  int off = block_offs[blockIdx.x] + incl - my;
  for (int w = 0; w < warp; ++w) off += s_w[w];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (!occ[j]) continue;
    const long long c = base + j;
    if (off < cap) {
      const int cnt = count[c];
      min_idx[off] = (long long)(unsigned)(best[c] & 0xffffffffull);
      if (centroid)
        for (int a = 0; a < 3; ++a) centroid[3 * (size_t)off + a] = vox_centroid(sums, c, a, cnt);
      if (grid_idx) {
        const int iz = (int)(c % v.M);
        const long long r = c / v.M;
        grid_idx[3 * (size_t)off] = (int)(r / v.M);
        grid_idx[3 * (size_t)off + 1] = (int)(r % v.M);
        grid_idx[3 * (size_t)off + 2] = iz;
      }
    }
    ++off;
  }
}

static inline long long vox_cells(int M) { return (long long)M * M * M; }

extern "C" size_t spf_voxelize_workspace_bytes(int32_t cells_per_axis) {
  const long long G = vox_cells(cells_per_axis);
  const long long nb = (G + VOX_TILE - 1) / VOX_TILE;
  // count int32 [G] | sums int64 [G,3] | best u64 [G] | block_sums int32 [nb] | err int32 (padded)
  return (size_t)(G * 4 + 16 + G * 24 + G * 8 + nb * 4 + 64);
}

extern "C" int spf_voxelize_closest(const float* points, int32_t n, float min_x, float min_y, float min_z, float vox_size,
                                    int32_t cells_per_axis, int64_t* min_idx, float* centroid, int32_t* grid_idx,
                                    int32_t cap, int32_t* n_out /*[2]: voxels, points outside the table*/, void* workspace,
                                    size_t workspace_bytes, void* stream_) {
  if (!points || !min_idx || !n_out || !workspace || n < 0 || cells_per_axis <= 0 || !(vox_size > 0.0f))
    return SPF_ERR_INVALID;
  const long long G = vox_cells(cells_per_axis);
  if (G >= (1LL << 31)) return SPF_ERR_UNSUPPORTED;
  if (workspace_bytes < spf_voxelize_workspace_bytes(cells_per_axis)) return SPF_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream_;
  const long long nb = (G + VOX_TILE - 1) / VOX_TILE;
  uint8_t* w = (uint8_t*)workspace;
  int* count = (int*)w;                          w += (G * 4 + 15) / 16 * 16;
  long long* sums = (long long*)w;               w += G * 24;
  unsigned long long* best = (unsigned long long*)w; w += G * 8;
  int* block_sums = (int*)w;                     w += (nb * 4 + 15) / 16 * 16;
  int* err = (int*)w;
  VoxDev v{min_x, min_y, min_z, vox_size, cells_per_axis};
  SPF_CUDA(cudaMemsetAsync(count, 0, (size_t)G * 4, st), "voxelize memset count");
  SPF_CUDA(cudaMemsetAsync(sums, 0, (size_t)G * 24, st), "voxelize memset sums");
  SPF_CUDA(cudaMemsetAsync(best, 0xff, (size_t)G * 8, st), "voxelize memset best");
  SPF_CUDA(cudaMemsetAsync(err, 0, 4, st), "voxelize memset err");
  if (n > 0) {
    k_vox_accum<<<(n + 255) / 256, 256, 0, st>>>(v, points, n, count, sums, err);
    SPF_CHECK_LAUNCH("k_vox_accum");
    k_vox_select<<<(n + 255) / 256, 256, 0, st>>>(v, points, n, count, sums, best);
    SPF_CHECK_LAUNCH("k_vox_select");
  }
  k_vox_count<<<(unsigned)nb, 256, 0, st>>>(count, G, block_sums);
  SPF_CHECK_LAUNCH("k_vox_count");
  k_vox_scan<<<1, 1024, 0, st>>>(block_sums, (int)nb, n_out);
  SPF_CHECK_LAUNCH("k_vox_scan");
  k_vox_emit<<<(unsigned)nb, 256, 0, st>>>(v, count, sums, best, G, block_sums, cap, (long long*)min_idx, centroid, grid_idx);
  SPF_CHECK_LAUNCH("k_vox_emit");
  SPF_CUDA(cudaMemcpyAsync(n_out + 1, err, 4, cudaMemcpyDeviceToDevice, st), "voxelize err");
  return SPF_OK;
}
