// mlp_tc2.cu -- second-generation bf16 tensor-core field kernels on the CTA-pair tile engine (tile_engine.cuh):
// tcgen05.mma.cta_group::2 (M = 256 over two CTAs), two row tiles in flight per CTA, weights through a TMA-fed ring,
// warp-specialised producer / MMA issuer / epilogue.  Same arithmetic, same saved-tensor layouts and same C ABI as
// mlp_tc.cu (whose kernels remain the single-CTA fallback for the radiance head and the self test).
//
// Row mapping: super-tile st (512 pair rows) -> 128-row tiles  tile = 4*st + 2*t + rank  (t = 0/1: tile X/Y of the CTA,
// rank = CTA rank in the pair); compact pair row = tile*128 + row, exactly as in mlp_tc.cu, so every saved tensor is
// interchangeable between the two generations.
#include <stdlib.h>
#include "tile_engine.cuh"

using namespace eng;

#define LEAKY 0.01f

__device__ __forceinline__ float rbf_w2(float dx, float dy, float dz, float rbf) {
  float dist = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
  float tq = dist * rbf;
  return __expf(-(tq * tq));
}

// ------------------------------------------------------------------------------------------------
// geometry field: gather -> 4 forward layers (-> d sdf / d input chain: 4 more) -> neighbour interpolation
// ------------------------------------------------------------------------------------------------
template <bool WITH_J>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
k_sdf_tc2(spf_geo_weights_tc W, const int* __restrict__ list, const int* __restrict__ count, const float* __restrict__ x,
          const int* __restrict__ pidx, const float* __restrict__ pts, const float* __restrict__ feat_g, float rbf,
          float* __restrict__ sdf, float* __restrict__ grad, float* __restrict__ jw) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Bars B = carve_bars(smem);
  float* s_part = reinterpret_cast<float*>(smem + OFF_PART);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int V = *count;
  const int ntiles = (V + 15) / 16;
  const int nsuper = (ntiles + 3) / 4;
  const int ncl = gridDim.x >> 1, cid = blockIdx.x >> 1;
  const int n_iter = cid < nsuper ? (nsuper - cid + ncl - 1) / ncl : 0;

  Chain ch;
  ch.L[0] = {W.w1p, 1, 3, 256};
  ch.L[1] = {W.w2p, 4, 16, 256};
  ch.L[2] = {W.w3p, 4, 16, 256};
  ch.L[3] = {W.w4p, 4, 16, 256};
  ch.L[4] = {W.w4tp, 4, 16, 256};
  ch.L[5] = {W.w3tp, 4, 16, 256};
  ch.L[6] = {W.w2tp, 4, 16, 256};
  ch.L[7] = {W.w1tp, 4, 16, 48};
  ch.n = WITH_J ? 8 : 4;

  const uint32_t tmem = setup(smem, B);

  if (warp == 8) {
    if (lane == 0) producer_loop(ch, n_iter, rank, smem, B);
  } else if (warp == 9) {
    if (lane == 0) {
      if (rank == 0) mma_loop(ch, n_iter, smem, B, tmem);
      else relay_loop(ch, n_iter, B);
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps
    const int row = 32 * (warp & 3) + lane;   // TMEM lane == pair row of the tile
    const int half = warp >> 2;               // which 128 accumulator columns this thread reads
    const uint32_t t_lane = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t acc_par = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int st = cid + it * ncl;
      int slot[2];
      float wgt[2];
      uint32_t bits[2][WITH_J ? 16 : 1];
      float dot[2] = {0.0f, 0.0f};
      // ---------------- gather: A0 = [g (32) | x_pi hi (3) | x_pi lo (3) | 0 ...] as bf16, K = 48
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        uint8_t* sA = smem + OFF_A + t * A_BYTES;
        const int tile = 4 * st + 2 * t + (int)rank;
        const int li = tile * 16 + (row >> 3);
        const int sl = li < V ? list[li] : -1;
        const int p = sl >= 0 ? pidx[(size_t)sl * 8 + (row & 7)] : -1;
        float xp[3] = {0.f, 0.f, 0.f}, w = 0.f;
        if (p >= 0) {
#pragma unroll
          for (int a = 0; a < 3; ++a) xp[a] = x[3 * (size_t)sl + a] - pts[3 * (size_t)p + a];
          w = rbf_w2(xp[0], xp[1], xp[2], rbf);
        }
        slot[t] = sl;
        wgt[t] = w;
        if (half == 0) {
          const float4* src = reinterpret_cast<const float4*>(feat_g + (size_t)(p >= 0 ? p : 0) * 32);
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            float4 a = p >= 0 ? src[2 * q] : make_float4(0, 0, 0, 0), b = p >= 0 ? src[2 * q + 1] : make_float4(0, 0, 0, 0);
            uint4 u = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
            *reinterpret_cast<uint4*>(sA + sw128_off(row, q)) = u;
          }
        } else {
          const float4* src = reinterpret_cast<const float4*>(feat_g + (size_t)(p >= 0 ? p : 0) * 32 + 24);
          float4 a = p >= 0 ? src[0] : make_float4(0, 0, 0, 0), b = p >= 0 ? src[1] : make_float4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(sA + sw128_off(row, 3)) =
              make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
          float hi[3], lo[3];
#pragma unroll
          for (int a3 = 0; a3 < 3; ++a3) {
            hi[a3] = __bfloat162float(__float2bfloat16_rn(xp[a3]));
            lo[a3] = xp[a3] - hi[a3];
          }
          *reinterpret_cast<uint4*>(sA + sw128_off(row, 4)) =
              make_uint4(pack_bf16(hi[0], hi[1]), pack_bf16(hi[2], lo[0]), pack_bf16(lo[1], lo[2]), 0u);
          *reinterpret_cast<uint4*>(sA + sw128_off(row, 5)) = make_uint4(0, 0, 0, 0);
        }
        signal_a_ready(B, t, rank);
      }
      // ---------------- forward chain
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const float* bias = l == 0 ? W.b1 : (l == 1 ? W.b2 : (l == 2 ? W.b3 : W.b4));
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          uint8_t* sA = smem + OFF_A + t * A_BYTES;
          wait_acc(B, t, acc_par);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c0 = half * 128 + q * 32;
            float v[32];
            tmem_ld32(t_lane + t * 256 + c0, v);
            float bb[32];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c0) + j4);
              bb[4 * j4] = b4.x; bb[4 * j4 + 1] = b4.y; bb[4 * j4 + 2] = b4.z; bb[4 * j4 + 3] = b4.w;
            }
            tmem_ld_wait();
            uint32_t b = 0;
            if (l < 3) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float tt = v[j] + bb[j];
                b |= (tt > 0.0f ? 1u : 0u) << j;
                v[j] = tt > 0.0f ? tt : LEAKY * tt;
              }
            } else {
              float vv[32];
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(W.v5 + c0) + j4);
                vv[4 * j4] = b4.x; vv[4 * j4 + 1] = b4.y; vv[4 * j4 + 2] = b4.z; vv[4 * j4 + 3] = b4.w;
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float tt = v[j] + bb[j];
                const bool pos = tt > 0.0f;
                b |= (pos ? 1u : 0u) << j;
                tt = pos ? tt : LEAKY * tt;
                dot[t] = fmaf(tt, vv[j], dot[t]);
                v[j] = vv[j] * (pos ? 1.0f : LEAKY);   // g4 = v5 * lrelu'(z4)
              }
            }
            if (WITH_J) bits[t][l * 4 + q] = b;
            if (l < 3 || WITH_J) store_a32(sA, row, c0, v);
          }
          if (l == 3) s_part[t * 256 + half * 128 + row] = dot[t];
          if (l < 3 || WITH_J) signal_a_ready(B, t, rank);
        }
      }
      if (WITH_J) {
        // ---------------- d sdf / d input chain: g_l = (g_{l+1} @ W_{l+1}) * lrelu'(z_l), J = g1 @ W1
#pragma unroll
        for (int l = 2; l >= 0; --l) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            uint8_t* sA = smem + OFF_A + t * A_BYTES;
            wait_acc(B, t, acc_par);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int c0 = half * 128 + q * 32;
              float v[32];
              tmem_ld32(t_lane + t * 256 + c0, v);
              tmem_ld_wait();
              const uint32_t b = bits[t][l * 4 + q];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= ((b >> j) & 1u) ? 1.0f : LEAKY;
              store_a32(sA, row, c0, v);
            }
            signal_a_ready(B, t, rank);
          }
        }
      } else {
        epi_bar();   // s_part visible to the other column half
      }
      // ---------------- Jacobian rows + neighbour interpolation
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int tile = 4 * st + 2 * t + (int)rank;
        const int sl = slot[t];
        float norm = wgt[t];
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) norm += __shfl_xor_sync(SPF_FULL, norm, o);
        const float wn = sl >= 0 ? wgt[t] / norm : 0.0f;
        if (WITH_J) {
          wait_acc(B, t, acc_par);
          if (half == 0) {
            float v[32];
            tmem_ld32(t_lane + t * 256, v);        // d sdf_k / d latent (32 columns)
            tmem_ld_wait();
            if (jw && tile < ntiles) {
              float4* dst = reinterpret_cast<float4*>(jw + ((size_t)tile * 128 + row) * 32);
#pragma unroll
              for (int q = 0; q < 8; ++q) dst[q] = make_float4(wn * v[4 * q], wn * v[4 * q + 1], wn * v[4 * q + 2], wn * v[4 * q + 3]);
            }
            tmem_ld32(t_lane + t * 256 + 32, v);   // columns 32..34 = d sdf_k / d (x - p)
            tmem_ld_wait();
            if (grad) {
              float g0 = wn * v[0], g1 = wn * v[1], g2 = wn * v[2];
#pragma unroll
              for (int o = 1; o < 8; o <<= 1) {
                g0 += __shfl_xor_sync(SPF_FULL, g0, o); g1 += __shfl_xor_sync(SPF_FULL, g1, o); g2 += __shfl_xor_sync(SPF_FULL, g2, o);
              }
              if (sl >= 0 && (row & 7) == 0) { grad[3 * (size_t)sl] = g0; grad[3 * (size_t)sl + 1] = g1; grad[3 * (size_t)sl + 2] = g2; }
            }
          }
        }
        if (half == 0) {
          float agg = wn * (s_part[t * 256 + row] + s_part[t * 256 + 128 + row] + W.c5);
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) agg += __shfl_xor_sync(SPF_FULL, agg, o);
          if (sl >= 0 && (row & 7) == 0) sdf[sl] = agg;
        }
      }
      tc_fence_before();
      epi_bar();   // s_part reads done before the next iteration overwrites it; TMEM reads ordered before the next gather's arrive
    }
  }
  teardown(tmem);
}

static bool use_gen1() {
  const char* e = getenv("SPF_TC_GEN1");
  return e && e[0] == '1';
}

extern "C" int spf_sdf_fwd_tc_gen1(const spf_geo_weights_tc* W, const int32_t* list, const int32_t* count, int64_t n_max,
                                   const float* x, const int32_t* pidx, int32_t K, const float* pts, const float* feat_g,
                                   float rbf, float* sdf, float* grad, float* jw, void* stream_);

extern "C" int spf_sdf_fwd_tc(const spf_geo_weights_tc* W, const int32_t* list, const int32_t* count, int64_t n_max,
                              const float* x, const int32_t* pidx, int32_t K, const float* pts, const float* feat_g,
                              float rbf, float* sdf, float* grad, float* jw, void* stream_) {
  if (use_gen1()) return spf_sdf_fwd_tc_gen1(W, list, count, n_max, x, pidx, K, pts, feat_g, rbf, sdf, grad, jw, stream_);
  if (!W || !list || !count || !x || !pidx || !pts || !feat_g || !sdf) return SPF_ERR_INVALID;
  if (K != 8) return SPF_ERR_UNSUPPORTED;
  if (n_max <= 0) return SPF_OK;
  cudaStream_t st = (cudaStream_t)stream_;
  const bool with_j = grad || jw;
  const int64_t supers = (n_max + 63) / 64;               // 64 slots = 512 pair rows per cluster iteration
  const int max_cl = spf_num_sms() / 2;
  const int grid = 2 * (int)(supers < max_cl ? supers : max_cl);
  if (with_j) {
    SPF_CUDA(cudaFuncSetAttribute(k_sdf_tc2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES), "sdf_tc2 attr");
    k_sdf_tc2<true><<<grid, THREADS, SMEM_BYTES, st>>>(*W, list, count, x, pidx, pts, feat_g, rbf, sdf, grad, jw);
  } else {
    SPF_CUDA(cudaFuncSetAttribute(k_sdf_tc2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES), "sdf_tc2 attr");
    k_sdf_tc2<false><<<grid, THREADS, SMEM_BYTES, st>>>(*W, list, count, x, pidx, pts, feat_g, rbf, sdf, grad, jw);
  }
  SPF_CHECK_LAUNCH("k_sdf_tc2");
  return SPF_OK;
}
