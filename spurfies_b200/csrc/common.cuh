// common.cuh -- shared helpers for the sm_100a kernels behind include/spurfies_b200.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/spurfies_b200.h"

#define SPF_FULL 0xffffffffu

extern "C" void spf_set_cuda_error(cudaError_t e, const char* where);

#define SPF_CHECK_LAUNCH(where)                                 \
  do {                                                          \
    cudaError_t _e = cudaGetLastError();                        \
    if (_e != cudaSuccess) {                                    \
      spf_set_cuda_error(_e, where);                            \
      return SPF_ERR_CUDA;                                      \
    }                                                           \
  } while (0)

#define SPF_CUDA(call, where)                                   \
  do {                                                          \
    cudaError_t _e = (call);                                    \
    if (_e != cudaSuccess) {                                    \
      spf_set_cuda_error(_e, where);                            \
      return SPF_ERR_CUDA;                                      \
    }                                                           \
  } while (0)

static inline int spf_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SPF_FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(SPF_FULL, v, o));
  return v;
}
// inclusive scan across the warp
__device__ __forceinline__ float warp_scan_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(SPF_FULL, v, o);
    if (lane >= o) v += n;
  }
  return v;
}
__device__ __forceinline__ int warp_scan_incl_i(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(SPF_FULL, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// grid geometry passed by value to kernels
struct GridDev {
  float sx, sy, sz;     // shift
  float vx, vy, vz;     // voxel edge
  int dx, dy, dz;       // dims
  int kx, ky, kz;       // kernel size
  const int* __restrict__ cell_start;
  const float4* __restrict__ sorted;
  const uint8_t* __restrict__ hit;
  // search grid (optional): cubic cells of edge fcell >= query radius
  float fcell;
  int fdx, fdy, fdz;
  const int* __restrict__ fcell_start;
  const float4* __restrict__ fsorted;
  int use_search;       // set per launch: search grid present and radius <= fcell
};
static inline GridDev to_dev(const spf_grid* g) {
  GridDev d;
  d.sx = g->shift[0]; d.sy = g->shift[1]; d.sz = g->shift[2];
  d.vx = g->vsize[0]; d.vy = g->vsize[1]; d.vz = g->vsize[2];
  d.dx = g->dim[0]; d.dy = g->dim[1]; d.dz = g->dim[2];
  d.kx = g->ks[0]; d.ky = g->ks[1]; d.kz = g->ks[2];
  d.cell_start = g->cell_start;
  d.sorted = reinterpret_cast<const float4*>(g->sorted);
  d.hit = g->hit;
  d.fcell = g->search_cell;
  d.fdx = g->search_dim[0]; d.fdy = g->search_dim[1]; d.fdz = g->search_dim[2];
  d.fcell_start = g->search_cell_start;
  d.fsorted = reinterpret_cast<const float4*>(g->search_sorted);
  d.use_search = 0;
  return d;
}

// voxel of a position, reference arithmetic: (int)floor((p - shift) / vsize) in fp32, IEEE division
// (torch_knnquery/src/knnquery.cu:45-47).  Returns linear cell or -1 when outside.
__device__ __forceinline__ int voxel_of(const GridDev& g, float x, float y, float z, int& cx, int& cy, int& cz) {
  cx = (int)floorf(__fdiv_rn(__fsub_rn(x, g.sx), g.vx));
  cy = (int)floorf(__fdiv_rn(__fsub_rn(y, g.sy), g.vy));
  cz = (int)floorf(__fdiv_rn(__fsub_rn(z, g.sz), g.vz));
  if (cx < 0 || cx >= g.dx || cy < 0 || cy >= g.dy || cz < 0 || cz >= g.dz) return -1;
  return cx * (g.dy * g.dz) + cy * g.dz + cz;
}
