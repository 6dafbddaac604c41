// mlp_tc.cu -- bf16 tensor-core mode, single-CTA kernels: the split-K weight-gradient kernels (row-major operands via a
// cp.async ring; tile-layout operands via TMA bulk copies), weight packing and the tcgen05 building-block self test.
// The per-pair fields and the per-sample radiance head live in mlp_tc2.cu on the CTA-pair tile engine.
#include "common.cuh"
#include <cuda_fp16.h>
#include "umma.cuh"

using namespace tc;

#define TC_THREADS 256
#define TC_ROWS 128

// ------------------------------------------------------------------------------------------------
// weight gradient: dW[256][N] += dZ^T @ A, db[256] += column sums of dZ, over the rows written by the dgrad kernels
// (whole 128-row tiles; rows of invalid slots hold zeros), row count taken from the device-side slot count so the
// step needs no host synchronisation.  HBM-bound (each operand row is read once): split-K over persistent CTAs,
// 3-stage cp.async ring of 64-row tiles, both operands MN-major, two M=128 accumulators fill TMEM (512 columns),
// fp32 vector atomics merge the per-CTA partials.
// ------------------------------------------------------------------------------------------------
#define WG_STAGE_BYTES 65536
#define WG_STAGES 3
#define WG_SMEM (WG_STAGES * WG_STAGE_BYTES + 256)
#define WGM_THREADS (TC_THREADS + 32)   // k_wgrad_multi: 8 compute warps + 1 producer warp

__global__ void __launch_bounds__(TC_THREADS, 1)
k_wgrad_tc(const __nv_bfloat16* __restrict__ dz, const __nv_bfloat16* __restrict__ act, int lda, int N,
           const int* __restrict__ count, int rows_per_unit, int layout, float* __restrict__ dW, float* __restrict__ db) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar_s = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);  // [WG_STAGES] stage free
  uint64_t* bar_done = bar_s + WG_STAGES;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long rows = ((long long)(*count) * rows_per_unit + TC_ROWS - 1) / TC_ROWS * TC_ROWS;
  const int ntiles = (int)(rows / 64);
  if ((int)blockIdx.x >= ntiles) return;
  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; ++s) mbar_init(bar_s + s, 1);
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t sbase = smem_u32(smem);
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nchunk_act = N / 8;  // 16-byte chunks per act row
  const uint32_t idesc = idesc_bf16_mn(128, N);
  float bsum = 0.0f;

  // operand addressing: row-major [rows, ld] or the tile engine's TILE layout (tile_engine.cuh: 128-row tiles, k-blocks
  // of 64 columns, 128-byte rows, 16-byte chunk c of row r at position c ^ (r & 7)) -- there a 64-row x 64-column
  // panel is one contiguous 8 KB run that already has the swizzle this kernel's MN-major descriptors expect
  const int act_nkb = (lda + 63) >> 6;
  auto src_chunk = [&](const __nv_bfloat16* base, bool tiled, long long row, int chunk, int ld, int nkb) -> const void* {
    if (!tiled) return base + row * ld + chunk * 8;
    const long long off = (row >> 7) * ((long long)nkb * 16384) + (long long)(chunk >> 3) * 16384 + (row & 127) * 128 +
                          (((chunk & 7) ^ (int)(row & 7)) << 4);
    return reinterpret_cast<const uint8_t*>(base) + off;
  };
  auto load_tile = [&](int it) {
    const int st = it % WG_STAGES;
    const long long r0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * 64;
    const uint32_t sdz = sbase + st * WG_STAGE_BYTES, sact = sdz + 32768;
    for (int e = tid; e < 64 * 32; e += TC_THREADS) {
      const int k = e >> 5, mc = e & 31;
      cp_async16(sdz + (mc >> 3) * 8192 + k * 128 + (((mc & 7) ^ (k & 7)) << 4), src_chunk(dz, layout & 1, r0 + k, mc, 256, 4));
    }
    for (int e = tid; e < 64 * nchunk_act; e += TC_THREADS) {
      const int k = e / nchunk_act, nc = e - k * nchunk_act;
      cp_async16(sact + (nc >> 3) * 8192 + k * 128 + (((nc & 7) ^ (k & 7)) << 4), src_chunk(act, layout & 2, r0 + k, nc, lda, act_nkb));
    }
  };

  for (int it = 0; it < WG_STAGES - 1; ++it) {
    if (it < my_tiles) load_tile(it);
    cp_async_commit();
  }
  for (int it = 0; it < my_tiles; ++it) {
    const int st = it % WG_STAGES;
    cp_async_wait<WG_STAGES - 2>();   // tile `it` has landed (only the most recent group may still be in flight)
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();                  // ... for every thread; also: everyone is done reading tile it-1
    const uint32_t sdz = sbase + st * WG_STAGE_BYTES, sact = sdz + 32768;
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t bd = smem_desc_mn_sw128(sact + ks * 2048, 8192);
        mma_bf16(tmem, smem_desc_mn_sw128(sdz + ks * 2048, 8192), bd, idesc, (it | ks) != 0);
        mma_bf16(tmem + 256, smem_desc_mn_sw128(sdz + 16384 + ks * 2048, 8192), bd, idesc, (it | ks) != 0);
      }
      mma_commit(bar_s + st);
    }
    // refill the stage tile it-1 used (tile it+2 maps to it) once its MMAs have retired
    const int nx = it + WG_STAGES - 1;
    if (nx < my_tiles) {
      if (it >= 1) mbar_wait(bar_s + (nx % WG_STAGES), (uint32_t)(((it - 1) / WG_STAGES) & 1));
      load_tile(nx);
    }
    cp_async_commit();
    // bias gradient: thread = output column, column sum over this tile's 64 rows (read from the swizzled tile)
    {
      const uint8_t* pdz = smem + st * WG_STAGE_BYTES + (tid >> 6) * 8192;
      const int c = (tid & 63) >> 3, e = tid & 7;
#pragma unroll 8
      for (int k = 0; k < 64; ++k) {
        const __nv_bfloat16 v = *reinterpret_cast<const __nv_bfloat16*>(pdz + k * 128 + ((c ^ (k & 7)) << 4) + e * 2);
        bsum += __bfloat162float(v);
      }
    }
  }
  if (tid == 0) mma_commit(bar_done);
  mbar_wait(bar_done, 0);
  tc_fence_after();
  if (db) atomicAdd(db + tid, bsum);
  // epilogue: thread = accumulator lane; warps 0-3 drain out rows 0..127, warps 4-7 rows 128..255
  {
    const int out_row = 128 * (warp >> 2) + 32 * (warp & 3) + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (warp >> 2) * 256;
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tmem_ld32(t_lane + c0, v);
      tmem_ld_wait();
      float* dst = dW + (size_t)out_row * N + c0;
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        if (c0 + 4 * j4 < N) atomicAdd(reinterpret_cast<float4*>(dst) + j4, make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// Same product for operands in the tile layout (layout == 3): a 64-row x 64-column panel is one contiguous, already
// swizzled 8 KB run, so the stages are filled by the TMA engine (8 KB bulk copies on a "full" mbarrier) instead of 4096
// cp.async per tile, and all 256 threads are free for the bias-gradient column sums.  Warp 0 issues the MMAs (one
// elected lane), thread 32 refills a stage once its MMAs have retired and all 8 warps have read it ("empty" mbarrier).
__global__ void __launch_bounds__(TC_THREADS, 1)
k_wgrad_tiled(const uint8_t* __restrict__ dz, const uint8_t* __restrict__ act, int act_nkb, int N,
              const int* __restrict__ count, int rows_per_unit, float* __restrict__ dW, float* __restrict__ db) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* bar_empty = bar_full + WG_STAGES;
  uint64_t* bar_done = bar_empty + WG_STAGES;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long rows = ((long long)(*count) * rows_per_unit + TC_ROWS - 1) / TC_ROWS * TC_ROWS;
  const int ntiles = (int)(rows / 64);
  if ((int)blockIdx.x >= ntiles) return;
  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 9); }
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t sbase = smem_u32(smem);
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int act_panels = (N + 63) >> 6;
  const bool conv_dz = false, conv_act = false;   // single-job entry point: both operands bf16
  const uint32_t idesc = idesc_bf16_mn(128, N);
  float bsum = 0.0f;

  auto load_tile = [&](int it) {   // one thread
    const int st = it % WG_STAGES;
    const long long j = (long long)blockIdx.x + (long long)it * gridDim.x;   // 64-row tile
    uint8_t* sdz = smem + st * WG_STAGE_BYTES;
    uint8_t* sact = sdz + 32768;
    mbar_expect_tx(bar_full + st, (uint32_t)(4 + act_panels) * 8192u);
    const uint8_t* gdz = dz + (j >> 1) * (4ll * 16384) + (j & 1) * 8192;
    const uint8_t* gact = act + (j >> 1) * ((long long)act_nkb * 16384) + (j & 1) * 8192;
    for (int kb = 0; kb < 4; ++kb) bulk_g2s(sdz + kb * 8192, gdz + kb * 16384, 8192, bar_full + st);
    for (int kb = 0; kb < act_panels; ++kb) bulk_g2s(sact + kb * 8192, gact + kb * 16384, 8192, bar_full + st);
  };

  if (tid == 32)
    for (int it = 0; it < WG_STAGES && it < my_tiles; ++it) load_tile(it);
  for (int it = 0; it < my_tiles; ++it) {
    const int st = it % WG_STAGES;
    const uint32_t ph = (uint32_t)((it / WG_STAGES) & 1);
    mbar_wait(bar_full + st, ph);
    const uint32_t sdz = sbase + st * WG_STAGE_BYTES, sact = sdz + 32768;
    if (conv_dz || conv_act) {
#pragma unroll 1
      for (int op = 0; op < 2; ++op) {
        if (!(op == 0 ? conv_dz : conv_act)) continue;
        uint8_t* base = smem + st * WG_STAGE_BYTES + op * 32768;
        const int nchunk = (op == 0 ? 4 : act_panels) * 512;   // 16-byte chunks of the operand's stage region
        // four chunks per thread and pass, all loads issued before the first conversion (nchunk is a multiple of 1024)
#pragma unroll 1
        for (int c0 = tid; c0 < nchunk; c0 += 4 * TC_THREADS) {
          uint4 u[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) u[q] = *reinterpret_cast<uint4*>(base + (c0 + q * TC_THREADS) * 16);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t* w = reinterpret_cast<uint32_t*>(&u[q]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
              w[e] = pack_bf16(f.x, f.y);
            }
            *reinterpret_cast<uint4*>(base + (c0 + q * TC_THREADS) * 16) = u[q];
          }
        }
      }
      fence_proxy_async();   // generic-proxy writes -> visible to the MMAs (async proxy) and ordered before the refill
      __syncthreads();
    }
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t bd = smem_desc_mn_sw128(sact + ks * 2048, 8192);
          mma_bf16(tmem, smem_desc_mn_sw128(sdz + ks * 2048, 8192), bd, idesc, (it | ks) != 0);
          mma_bf16(tmem + 256, smem_desc_mn_sw128(sdz + 16384 + ks * 2048, 8192), bd, idesc, (it | ks) != 0);
        }
        mma_commit(bar_empty + st);
      }
      __syncwarp();
    }
    // bias gradient: thread = output column, column sum over this tile's 64 rows (read from the swizzled tile)
    {
      const uint8_t* pdz = smem + st * WG_STAGE_BYTES + (tid >> 6) * 8192;
      const int c = (tid & 63) >> 3, e = tid & 7;
#pragma unroll 8
      for (int k = 0; k < 64; ++k) {
        const __nv_bfloat16 v = *reinterpret_cast<const __nv_bfloat16*>(pdz + k * 128 + ((c ^ (k & 7)) << 4) + e * 2);
        bsum += __bfloat162float(v);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_local(bar_empty + st);
    if (tid == 32 && it + WG_STAGES < my_tiles) {
      mbar_wait(bar_empty + st, ph);   // MMAs of tile `it` retired and every warp has read the stage
      load_tile(it + WG_STAGES);
    }
  }
  if (warp == 0) {
    if (elect_one()) mma_commit(bar_done);
    __syncwarp();
  }
  mbar_wait(bar_done, 0);
  tc_fence_after();
  if (db) atomicAdd(db + tid, bsum);
  {
    const int out_row = 128 * (warp >> 2) + 32 * (warp & 3) + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (warp >> 2) * 256;
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tmem_ld32(t_lane + c0, v);
      tmem_ld_wait();
      float* dst = dW + (size_t)out_row * N + c0;
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        if (c0 + 4 * j4 < N) atomicAdd(reinterpret_cast<float4*>(dst) + j4, make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// Several weight-gradient products over the SAME rows in ONE launch (all operands in the tile layout): the persistent CTAs
// are partitioned among the jobs in proportion to their bytes per row, so every CTA still streams one product but over a
// longer row range -- one launch / TMEM allocation / tail instead of one per layer, and the fp32 atomic merge traffic
// (256 x N floats per CTA) is paid once per CTA instead of once per CTA and layer.  Body = k_wgrad_tiled.
struct WgJobDev { const uint8_t* dz; const uint8_t* act; float* dW; float* db; int act_nkb, N, cta0, ncta, fmt; };
struct WgJobsDev { WgJobDev j[SPF_WGRAD_MAX_JOBS]; int n; };

__global__ void __launch_bounds__(WGM_THREADS, 1)
k_wgrad_multi(WgJobsDev jobs, const int* __restrict__ count, int rows_per_unit, const float* __restrict__ gscale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* bar_empty = bar_full + WG_STAGES;
  uint64_t* bar_done = bar_empty + WG_STAGES;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int q = 0;
  while (q + 1 < jobs.n && (int)blockIdx.x >= jobs.j[q + 1].cta0) ++q;
  const WgJobDev J = jobs.j[q];
  const int bid = (int)blockIdx.x - J.cta0, nb = J.ncta;
  const long long rows = ((long long)(*count) * rows_per_unit + TC_ROWS - 1) / TC_ROWS * TC_ROWS;
  const int ntiles = (int)(rows / 64);
  if (bid >= ntiles) return;
  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 9); }
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t sbase = smem_u32(smem);
  const int my_tiles = (ntiles - bid + nb - 1) / nb;
  const int N = J.N;
  const int act_panels = (N + 63) >> 6;
  // tcgen05.mma kind::f16 wants A and B in the SAME 16-bit format (a bf16 x fp16 instruction descriptor is an illegal
  // instruction on the B200).  Gradient tiles are bf16 and the saved forward activations fp16, so the fp16 operand of
  // a stage is converted to bf16 IN PLACE in shared memory (elementwise, so the swizzle does not matter) by all 256
  // threads before the MMAs of that stage are issued: ~110 instructions per thread per 64-row tile, against a
  // ~3000-cycle HBM budget per tile.  Rounding fp16 -> bf16 here is a plain 2^-9 rounding of a wgrad operand; the
  // LeakyReLU sign decisions were taken in the forward pass on the fp16 values.
  // fmt 0: both fp16 (the training step: scaled-fp16 gradient tiles x fp16 saved activations) and fmt 3: both bf16 go
  // straight to the tensor cores; a mixed pair has its fp16 operand converted to bf16 in shared memory first
  const bool both_f16 = (J.fmt & 3) == 0;
  const uint32_t idesc = idesc_f16k_mn(128, N, both_f16 ? FMT_F16 : FMT_BF16);
  const bool conv_dz = !both_f16 && !(J.fmt & 1), conv_act = !both_f16 && !(J.fmt & 2);
  const bool dz_f16 = both_f16;             // format of dZ in shared memory when the bias sums read it
  const float inv_s = gscale ? gscale[1] : 1.0f;   // the gradient tiles carry the step's power-of-two scale S
  float bsum = 0.0f;

  auto load_tile = [&](int it) {   // one thread
    const int st = it % WG_STAGES;
    const long long j = (long long)bid + (long long)it * nb;   // 64-row tile
    uint8_t* sdz = smem + st * WG_STAGE_BYTES;
    uint8_t* sact = sdz + 32768;
    mbar_expect_tx(bar_full + st, (uint32_t)(4 + act_panels) * 8192u);
    const uint8_t* gdz = J.dz + (j >> 1) * (4ll * 16384) + (j & 1) * 8192;
    const uint8_t* gact = J.act + (j >> 1) * ((long long)J.act_nkb * 16384) + (j & 1) * 8192;
    for (int kb = 0; kb < 4; ++kb) bulk_g2s(sdz + kb * 8192, gdz + kb * 16384, 8192, bar_full + st);
    for (int kb = 0; kb < act_panels; ++kb) bulk_g2s(sact + kb * 8192, gact + kb * 16384, 8192, bar_full + st);
  };

  if (warp == 8) {
    // producer warp: refills a stage as soon as its MMAs have retired and all 8 compute warps have left it.  Decoupled
    // from the compute warps (they synchronise among themselves on a named barrier), so waiting for an MMA to retire
    // never holds up the conversion / bias sums of the following tiles.
    if (lane == 0)
      for (int it = 0; it < my_tiles; ++it) {
        if (it >= WG_STAGES) mbar_wait(bar_empty + it % WG_STAGES, (uint32_t)(((it / WG_STAGES) - 1) & 1));
        load_tile(it);
      }
  } else
  for (int it = 0; it < my_tiles; ++it) {
    const int st = it % WG_STAGES;
    const uint32_t ph = (uint32_t)((it / WG_STAGES) & 1);
    mbar_wait(bar_full + st, ph);
    const uint32_t sdz = sbase + st * WG_STAGE_BYTES, sact = sdz + 32768;
    if (conv_dz || conv_act) {
#pragma unroll 1
      for (int op = 0; op < 2; ++op) {
        if (!(op == 0 ? conv_dz : conv_act)) continue;
        uint8_t* base = smem + st * WG_STAGE_BYTES + op * 32768;
        const int nchunk = (op == 0 ? 4 : act_panels) * 512;   // 16-byte chunks of the operand's stage region
        // four chunks per thread and pass, all loads issued before the first conversion (nchunk is a multiple of 1024)
#pragma unroll 1
        for (int c0 = tid; c0 < nchunk; c0 += 4 * TC_THREADS) {
          uint4 u[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) u[q] = *reinterpret_cast<uint4*>(base + (c0 + q * TC_THREADS) * 16);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t* w = reinterpret_cast<uint32_t*>(&u[q]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
              w[e] = pack_bf16(f.x, f.y);
            }
            *reinterpret_cast<uint4*>(base + (c0 + q * TC_THREADS) * 16) = u[q];
          }
        }
      }
      fence_proxy_async();   // generic-proxy writes -> visible to the MMAs (async proxy) and ordered before the refill
      named_bar_sync(1, TC_THREADS);
    }
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t bd = smem_desc_mn_sw128(sact + ks * 2048, 8192);
          mma_bf16(tmem, smem_desc_mn_sw128(sdz + ks * 2048, 8192), bd, idesc, (it | ks) != 0);
          mma_bf16(tmem + 256, smem_desc_mn_sw128(sdz + 16384 + ks * 2048, 8192), bd, idesc, (it | ks) != 0);
        }
        mma_commit(bar_empty + st);
      }
      __syncwarp();
    }
    if (J.db) {   // bias gradient: thread = output column, column sum over this tile's 64 rows (read from the swizzled tile)
      const uint8_t* pdz = smem + st * WG_STAGE_BYTES + (tid >> 6) * 8192;
      const int c = (tid & 63) >> 3, e = tid & 7;
#pragma unroll 8
      for (int k = 0; k < 64; ++k) {
        const unsigned short raw = *reinterpret_cast<const unsigned short*>(pdz + k * 128 + ((c ^ (k & 7)) << 4) + e * 2);
        bsum += dz_f16 ? __half2float(__ushort_as_half(raw)) : __uint_as_float((unsigned)raw << 16);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_local(bar_empty + st);
  }
  if (warp == 0) {
    if (elect_one()) mma_commit(bar_done);
    __syncwarp();
  }
  mbar_wait(bar_done, 0);
  tc_fence_after();
  if (J.db && warp < 8) atomicAdd(J.db + tid, inv_s * bsum);
  if (warp < 8) {
    const int out_row = 128 * (warp >> 2) + 32 * (warp & 3) + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (warp >> 2) * 256;
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tmem_ld32(t_lane + c0, v);
      tmem_ld_wait();
      float* dst = J.dW + (size_t)out_row * N + c0;
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        if (c0 + 4 * j4 < N) atomicAdd(reinterpret_cast<float4*>(dst) + j4, make_float4(inv_s * v[4 * j4], inv_s * v[4 * j4 + 1], inv_s * v[4 * j4 + 2], inv_s * v[4 * j4 + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

extern "C" int spf_wgrad_tc_multi(const spf_wgrad_job* jobs, int32_t n_jobs, const int32_t* count, int32_t rows_per_unit,
                                  int64_t n_max, const float* gscale, void* stream_) {
  if (!jobs || !count || n_jobs < 1 || n_jobs > SPF_WGRAD_MAX_JOBS || rows_per_unit < 1) return SPF_ERR_INVALID;
  if (n_max <= 0) return SPF_OK;
  WgJobsDev d;
  d.n = n_jobs;
  const int sms = spf_num_sms();
  if (sms < n_jobs) return SPF_ERR_UNSUPPORTED;
  long long wsum = 0, w[SPF_WGRAD_MAX_JOBS];
  for (int q = 0; q < n_jobs; ++q) {
    const spf_wgrad_job& J = jobs[q];
    if (!J.dz || !J.act || !J.dW) return SPF_ERR_INVALID;
    if (J.N % 16 || J.N < 16 || J.N > 256 || J.lda < J.N || J.lda % 64) return SPF_ERR_UNSUPPORTED;
    w[q] = 512 + (long long)(J.lda >> 6) * 128;   // bytes per row of this product
    wsum += w[q];
  }
  int used = 0;
  for (int q = 0; q < n_jobs; ++q) {
    int c = (int)((w[q] * sms) / wsum);
    if (c < 1) c = 1;
    if (q == n_jobs - 1 || used + c > sms - (n_jobs - 1 - q)) c = sms - (n_jobs - 1 - q) - used;
    const spf_wgrad_job& J = jobs[q];
    d.j[q] = {(const uint8_t*)J.dz, (const uint8_t*)J.act, J.dW, J.db, J.lda >> 6, J.N, used, c, J.fmt & 3};
    used += c;
  }
  SPF_CUDA(cudaFuncSetAttribute(k_wgrad_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM), "wgrad attr");
  k_wgrad_multi<<<used, WGM_THREADS, WG_SMEM, (cudaStream_t)stream_>>>(d, count, rows_per_unit, gscale);
  SPF_CHECK_LAUNCH("k_wgrad_multi");
  return SPF_OK;
}

extern "C" int spf_wgrad_tc(const void* dz, const void* act, int32_t lda, int32_t N, const int32_t* count,
                            int32_t rows_per_unit, int64_t n_max, int32_t layout, float* dW, float* db, void* stream_) {
  if (!dz || !act || !count || !dW) return SPF_ERR_INVALID;
  if (N % 16 || N < 16 || N > 256 || lda < N || lda % 8 || rows_per_unit < 1) return SPF_ERR_UNSUPPORTED;
  if (n_max <= 0) return SPF_OK;
  int64_t tiles = (n_max * rows_per_unit + 63) / 64;
  int grid = (int)(tiles < spf_num_sms() ? tiles : spf_num_sms());
  if (layout == 3) {
    SPF_CUDA(cudaFuncSetAttribute(k_wgrad_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM), "wgrad attr");
    k_wgrad_tiled<<<grid, TC_THREADS, WG_SMEM, (cudaStream_t)stream_>>>((const uint8_t*)dz, (const uint8_t*)act, (lda + 63) >> 6, N,
                                                                        count, rows_per_unit, dW, db);
  } else {
    SPF_CUDA(cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM), "wgrad attr");
    k_wgrad_tc<<<grid, TC_THREADS, WG_SMEM, (cudaStream_t)stream_>>>((const __nv_bfloat16*)dz, (const __nv_bfloat16*)act, lda,
                                                                    N, count, rows_per_unit, layout, dW, db);
  }
  SPF_CHECK_LAUNCH("k_wgrad_tc");
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
// weight image packing (fp32 [N][K] row-major, or its transpose, -> bf16 k-block-major 128B-swizzled image)
// thread = one 16-byte chunk of the image
// ------------------------------------------------------------------------------------------------
__global__ void k_pack_sw128(const float* __restrict__ W, int ld, int N, int K, int transpose, int n_pad, int nkb,
                             uint4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nkb * n_pad * 8) return;
  const int cs = i & 7, n = (i >> 3) % n_pad, kb = i / (8 * n_pad);
  const int c = cs ^ (n & 7);
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = kb * 64 + c * 8 + e;
    v[e] = (n < N && k < K) ? ((transpose & 1) ? W[(size_t)k * ld + n] : W[(size_t)n * ld + k]) : 0.0f;
  }
  out[i] = (transpose & SPF_PACK_F16)
               ? make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]))
               : make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

extern "C" int spf_pack_sw128(const float* W, int32_t ld, int32_t N, int32_t K, int32_t transpose, int32_t n_pad, void* out,
                              void* stream_) {
  if (!W || !out || N < 1 || K < 1 || n_pad < N || n_pad % 8) return SPF_ERR_INVALID;
  const int nkb = (K + 63) / 64;
  const int total = nkb * n_pad * 8;
  k_pack_sw128<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(W, ld, N, K, transpose, n_pad, nkb, (uint4*)out);
  SPF_CHECK_LAUNCH("k_pack_sw128");
  return SPF_OK;
}

// all weight images of a step in ONE launch (the trainable weights change every step): blockIdx.y = job
struct PackJobsDev { spf_pack_job j[SPF_PACK_MAX_JOBS]; };
__global__ void k_pack_sw128_batch(PackJobsDev jobs) {
  const spf_pack_job& J = jobs.j[blockIdx.y];
  const int nkb = (J.K + 63) / 64;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nkb * J.n_pad * 8) return;
  const int cs = i & 7, n = (i >> 3) % J.n_pad, kb = i / (8 * J.n_pad);
  const int c = cs ^ (n & 7);
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = kb * 64 + c * 8 + e;
    v[e] = (n < J.N && k < J.K) ? ((J.transpose & 1) ? J.W[(size_t)k * J.ld + n] : J.W[(size_t)n * J.ld + k]) : 0.0f;
  }
  reinterpret_cast<uint4*>(J.out)[i] =
      (J.transpose & SPF_PACK_F16)
          ? make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]))
          : make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

extern "C" int spf_pack_sw128_batch(const spf_pack_job* jobs, int32_t n_jobs, void* stream_) {
  if (!jobs || n_jobs < 1 || n_jobs > SPF_PACK_MAX_JOBS) return SPF_ERR_INVALID;
  PackJobsDev d;
  int max_total = 0;
  for (int q = 0; q < n_jobs; ++q) {
    const spf_pack_job& J = jobs[q];
    if (!J.W || !J.out || J.N < 1 || J.K < 1 || J.n_pad < J.N || J.n_pad % 8) return SPF_ERR_INVALID;
    d.j[q] = J;
    const int total = ((J.K + 63) / 64) * J.n_pad * 8;
    if (total > max_total) max_total = total;
  }
  dim3 grid((max_total + 255) / 256, n_jobs);
  k_pack_sw128_batch<<<grid, 256, 0, (cudaStream_t)stream_>>>(d);
  SPF_CHECK_LAUNCH("k_pack_sw128_batch");
  return SPF_OK;
}

// per-ray constant part of R.0 (pointneus_disent.py:100-107, embedder.py:10-36): zpe[r] = R.0.weight[:, :21] PE3(dir_r) + R.0.bias.
// A block stages the 256 x 21 weight slice in shared memory once (transposed: conflict-free reads) and walks ZPE_RAYS
// rays; thread = output unit.
#define ZPE_RAYS 16
__global__ void __launch_bounds__(256) k_head_zpe(const float* __restrict__ dirs, const float* __restrict__ W, int ld,
                                                  const float* __restrict__ bias, int R, float* __restrict__ zpe) {
  __shared__ float Ws[21 * 256];
  __shared__ float pe[ZPE_RAYS][24];
  const int tid = threadIdx.x;
  for (int e = tid; e < 256 * 21; e += 256) {
    const int row = e / 21, j = e - row * 21;
    Ws[j * 256 + row] = W[(size_t)row * ld + j];
  }
  const int r0 = blockIdx.x * ZPE_RAYS;
  for (int e = tid; e < ZPE_RAYS * 21; e += 256) {
    const int q = e / 21, k = e - q * 21, r = r0 + q;
    float v = 0.0f;
    if (r < R) {
      if (k < 3) v = dirs[3 * r + k];
      else {
        const int l = (k - 3) / 6, rem = (k - 3) % 6, a = rem % 3;
        const float arg = dirs[3 * r + a] * (float)(1 << l);
        v = rem < 3 ? sinf(arg) : cosf(arg);
      }
    }
    pe[q][k] = v;
  }
  __syncthreads();
  const float bv = bias[tid];
#pragma unroll 4
  for (int q = 0; q < ZPE_RAYS; ++q) {
    if (r0 + q >= R) break;
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < 21; ++j) acc = fmaf(pe[q][j], Ws[j * 256 + tid], acc);
    zpe[(size_t)(r0 + q) * 256 + tid] = acc + bv;
  }
}

extern "C" int spf_head_zpe(const float* dirs, const float* W, int32_t ld, const float* bias, int32_t R, float* zpe,
                            void* stream_) {
  if (!dirs || !W || !bias || !zpe || ld < 21) return SPF_ERR_INVALID;
  if (R <= 0) return SPF_OK;
  k_head_zpe<<<(R + ZPE_RAYS - 1) / ZPE_RAYS, 256, 0, (cudaStream_t)stream_>>>(dirs, W, ld, bias, R, zpe);
  SPF_CHECK_LAUNCH("k_head_zpe");
  return SPF_OK;
}

// -------------------------------------------------------------------------------------------------
// building-block self test: out[128][N] = A[128][K] (bf16 row-major, K multiple of 16, <= 256) @ Wp^T, where Wp is
// the packed (k-block major, SW128) image of W [N][K].  Exercises: generic-proxy writes of the A tile in the
// swizzled layout, bulk-copy weight load on an mbarrier, tcgen05.mma issue + commit, tcgen05.ld epilogue.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
k_tc_gemm_test(const __nv_bfloat16* __restrict__ A, const uint8_t* __restrict__ Wp, int N, int K, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                 // 4 k-blocks x 16 KB
  uint8_t* sW = smem + 65536;         // 4 k-blocks x N*128 B
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + 65536 + 131072);
  uint64_t* bar_m = bar_w + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_m + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = (K + 63) / 64;
  if (tid == 0) { mbar_init(bar_w, 1); mbar_init(bar_m, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(s_tmem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)nkb * N * 128;
    mbar_expect_tx(bar_w, bytes);
    for (int kb = 0; kb < nkb; ++kb) bulk_g2s(sW + (size_t)kb * N * 128, Wp + (size_t)kb * N * 128, N * 128, bar_w);
  }
  // A tile: thread -> (row, chunks)
  for (int e = tid; e < 128 * nkb * 8; e += 256) {
    int row = e / (nkb * 8), cc = e % (nkb * 8), kb = cc >> 3, ch = cc & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    int k0 = kb * 64 + ch * 8;
    if (k0 < K) v = *reinterpret_cast<const uint4*>(A + (size_t)row * K + k0);
    *reinterpret_cast<uint4*>(sA + kb * 16384 + sw128_off(row, ch)) = v;
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    mbar_wait(bar_w, 0);
    tc_fence_after();
    const uint32_t idesc = idesc_bf16(128, N);
    int acc = 0;
    for (int k = 0; k < K; k += 16) {
      int kb = k >> 6, ks = (k & 63) >> 4;
      uint64_t ad = smem_desc_sw128(smem_u32(sA) + kb * 16384 + ks * 32);
      uint64_t bd = smem_desc_sw128(smem_u32(sW) + kb * N * 128 + ks * 32);
      mma_bf16(tmem, ad, bd, idesc, acc);
      acc = 1;
    }
    mma_commit(bar_m);
  }
  mbar_wait(bar_m, 0);
  tc_fence_after();
  const int row = 32 * (warp & 3) + lane;
  const int half = warp >> 2;
  for (int c0 = half * 128; c0 < half * 128 + 128 && c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) out[(size_t)row * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

extern "C" int spf_tc_gemm_test(const void* A, const void* Wp, int32_t N, int32_t K, float* out, void* stream_) {
  if (!A || !Wp || !out) return SPF_ERR_INVALID;
  if (N % 16 || N < 16 || N > 256 || K % 16 || K < 16 || K > 256) return SPF_ERR_UNSUPPORTED;
  const int smem = 65536 + 131072 + 1024;
  SPF_CUDA(cudaFuncSetAttribute(k_tc_gemm_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "tc test attr");
  k_tc_gemm_test<<<1, 256, smem, (cudaStream_t)stream_>>>((const __nv_bfloat16*)A, (const uint8_t*)Wp, N, K, out);
  SPF_CHECK_LAUNCH("k_tc_gemm_test");
  return SPF_OK;
}
