// mlp_tc2.cu -- second-generation bf16 tensor-core field kernels on the CTA-pair tile engine (tile_engine.cuh):
// tcgen05.mma.cta_group::2 (M = 256 over two CTAs), two row tiles in flight per CTA, weights through a TMA-fed ring,
// warp-specialised producer / MMA issuer / epilogue.  Same arithmetic, same saved-tensor layouts and same C ABI as
// the first-generation single-CTA kernels they replace (profiles/r01b_*); mlp_tc.cu keeps the radiance head, the
// weight-gradient kernel and the self test.
//
// Row mapping: super-tile st (512 pair rows) -> 128-row tiles  tile = 4*st + 2*t + rank  (t = 0/1: tile X/Y of the CTA,
// rank = CTA rank in the pair); compact pair row = tile*128 + row.
#include <stdlib.h>
#include "tile_engine.cuh"

using namespace eng;

#define LEAKY 0.01f

// (L2 prefetches of the next tile's gather rows -- latent rows, points, the compact d_hbar lines, sign words -- were
// measured and rejected: no change within noise; the gathers already hit L2 and the kernels are epilogue-issue bound.)

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

  Chain& ch = *reinterpret_cast<Chain*>(smem + OFF_CHAIN);
  float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);   // b1..b4, v5
  if (tid == 0) {
    ch.L[0] = {W.w1p, 1, 3, 256, FMT_F16};      // forward layers: fp16 operands (umma.cuh)
    ch.L[1] = {W.w2p, 4, 16, 256, FMT_F16};
    ch.L[2] = {W.w3p, 4, 16, 256, FMT_F16};
    ch.L[3] = {W.w4p, 4, 16, 256, FMT_F16};
    ch.L[4] = {W.w4tp, 4, 16, 256, FMT_F16};    // d sdf / d input chain: the network's Jacobian (O(weights) magnitudes,
    ch.L[5] = {W.w3tp, 4, 16, 256, FMT_F16};    // no 1/R loss scaling), so fp16 is range-safe here too and keeps 3 more
    ch.L[6] = {W.w2tp, 4, 16, 256, FMT_F16};    // bits per layer than bf16
    ch.L[7] = {W.w1tp, 4, 16, 48, FMT_F16};
    ch.n = WITH_J ? 8 : 4;
  }
  for (int i = tid; i < 256; i += THREADS) {
    s_bias[i] = W.b1[i]; s_bias[256 + i] = W.b2[i]; s_bias[512 + i] = W.b3[i]; s_bias[768 + i] = W.b4[i];
    s_bias[1024 + i] = W.v5[i];
  }
  const uint32_t tmem = setup(smem, B);

  if (warp == WARP_PRODUCER) {
    if (lane == 0) producer_loop(ch, n_iter, rank, smem, B);
  } else if (warp == WARP_MMA) {
    if (rank == 0) mma_loop(ch, n_iter, smem, B, tmem);   // whole warp, one elected lane issues
    else if (lane == 0) relay_loop(ch, n_iter, B);
  } else {
    // ---------------------------------------------------------------- epilogue warps: group t owns tile t of this CTA
    const int t = warp >> 3;
    const int row = 32 * (warp & 3) + lane;   // TMEM lane == pair row of the tile
    const int half = (warp >> 2) & 1;         // which 128 accumulator columns this thread reads
    const uint32_t t_row = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + t * 256;
    const uint32_t t_acc = t_row + half * 128;
    uint8_t* sA = smem + OFF_A + t * A_BYTES;
    float* part = s_part + t * 256;
    uint32_t acc_par = 0;
    int sl_n = -1, p_n = -1;
    {
      const int li = (4 * cid + 2 * t + (int)rank) * 16 + (row >> 3);
      if (n_iter > 0 && li < V) { sl_n = list[li]; p_n = pidx[(size_t)sl_n * 8 + (row & 7)]; }
    }
    for (int it = 0; it < n_iter; ++it) {
      const int tile = 4 * (cid + it * ncl) + 2 * t + (int)rank;
      const int sl = sl_n, p = p_n;
      // ---------------- gather: A0 = [g (32) | x_pi hi (3) | x_pi lo (3) | 0 ...] as bf16, K = 48
      float xp[3] = {0.f, 0.f, 0.f}, w = 0.f;
      if (p >= 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) xp[a] = x[3 * (size_t)sl + a] - pts[3 * (size_t)p + a];
        w = rbf_w2(xp[0], xp[1], xp[2], rbf);
      }
      if (half == 0) {
        const float4* src = reinterpret_cast<const float4*>(feat_g + (size_t)(p >= 0 ? p : 0) * 32);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          float4 a = p >= 0 ? src[2 * q] : make_float4(0, 0, 0, 0), b = p >= 0 ? src[2 * q + 1] : make_float4(0, 0, 0, 0);
          uint4 u = make_uint4(pack_f16(a.x, a.y), pack_f16(a.z, a.w), pack_f16(b.x, b.y), pack_f16(b.z, b.w));
          *reinterpret_cast<uint4*>(sA + sw128_off(row, q)) = u;
        }
      } else {
        const float4* src = reinterpret_cast<const float4*>(feat_g + (size_t)(p >= 0 ? p : 0) * 32 + 24);
        float4 a = p >= 0 ? src[0] : make_float4(0, 0, 0, 0), b = p >= 0 ? src[1] : make_float4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sA + sw128_off(row, 3)) =
            make_uint4(pack_f16(a.x, a.y), pack_f16(a.z, a.w), pack_f16(b.x, b.y), pack_f16(b.z, b.w));
        float hi[3], lo[3];   // x - p as fp16 hi + lo (22 significant bits) against the same weight columns
#pragma unroll
        for (int a3 = 0; a3 < 3; ++a3) {
          hi[a3] = f16_round(xp[a3]);
          lo[a3] = xp[a3] - hi[a3];
        }
        *reinterpret_cast<uint4*>(sA + sw128_off(row, 4)) =
            make_uint4(pack_f16(hi[0], hi[1]), pack_f16(hi[2], lo[0]), pack_f16(lo[1], lo[2]), 0u);
        *reinterpret_cast<uint4*>(sA + sw128_off(row, 5)) = make_uint4(0, 0, 0, 0);
      }
      if ((tid & 255) == 0) TL(3, t, it);
      signal_a_ready(B, t, rank);
      // indices of the next super-tile: two dependent global loads taken off the next gather's critical path
      sl_n = -1; p_n = -1;
      if (it + 1 < n_iter) {
        const int li = (tile + 4 * ncl) * 16 + (row >> 3);
        if (li < V) { sl_n = list[li]; p_n = pidx[(size_t)sl_n * 8 + (row & 7)]; }
      }
      uint32_t bits[WITH_J ? 24 : 1];   // LeakyReLU sign bits of z1..z3 (local memory: the layer loop is a run-time loop)
      float dot = 0.0f;
      // ---------------- forward chain.  Epilogue cost matters as much as the MMAs here (K = 256 only): ~4.5 instructions
      // per element -- fp32 bias add, LeakyReLU as max(z, 0.01 z), bf16x2 pack, and the LeakyReLU sign bits taken from
      // the packed words (sign(h) == sign(z); 2 bits per word, accumulated with a shift + and-or).
#pragma unroll 1
      for (int l = 0; l < 4; ++l) {
        const float4* bias4 = reinterpret_cast<const float4*>(s_bias + l * 256 + half * 128);
        const float4* v54 = reinterpret_cast<const float4*>(s_bias + 1024 + half * 128);
        wait_acc(B, t, acc_par);
        if ((tid & 255) == 0) TL(0, t, l);
        float v[2][16];
        const int dbg = DBG_MODE;
        if (dbg != 1) tmem_ld16(t_acc, v[0]);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (dbg != 1) {
            tmem_ld_wait();
            if (c < 7) tmem_ld16(t_acc + (c + 1) * 16, v[(c + 1) & 1]);
          }
          float* vv = v[c & 1];
          uint32_t pk[8];
          uint32_t sb = 0;
          if (dbg == 2) { if (vv[0] == 123.456f) part[0] = vv[1]; continue; }
          if (l < 3) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 bq = bias4[c * 4 + q];
              const float z0 = vv[4 * q] + bq.x, z1 = vv[4 * q + 1] + bq.y, z2 = vv[4 * q + 2] + bq.z, z3 = vv[4 * q + 3] + bq.w;
              pk[2 * q] = leaky_f16x2(pack_f16(z0, z1), LEAKY_H2);      // sign(h) == sign(z): the sign bits below are exact
              pk[2 * q + 1] = leaky_f16x2(pack_f16(z2, z3), LEAKY_H2);
              if (WITH_J) {
                sb = (sb >> 2) | (pk[2 * q] & 0x80008000u);
                sb = (sb >> 2) | (pk[2 * q + 1] & 0x80008000u);
              }
            }
            // word i of this chunk: its low / high element's "negative" bit sits at bit 1 + 2 i / 17 + 2 i of sb
            if (WITH_J) bits[l * 8 + c] = sb;
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 bq = bias4[c * 4 + q];
              const float4 vq = v54[c * 4 + q];
              float z0 = vv[4 * q] + bq.x, z1 = vv[4 * q + 1] + bq.y, z2 = vv[4 * q + 2] + bq.z, z3 = vv[4 * q + 3] + bq.w;
              dot = fmaf(fmaxf(z0, LEAKY * z0), vq.x, dot); dot = fmaf(fmaxf(z1, LEAKY * z1), vq.y, dot);
              dot = fmaf(fmaxf(z2, LEAKY * z2), vq.z, dot); dot = fmaf(fmaxf(z3, LEAKY * z3), vq.w, dot);
              if (WITH_J) {   // g4 = v5 * lrelu'(z4)
                pk[2 * q] = pack_f16(z0 > 0.0f ? vq.x : LEAKY * vq.x, z1 > 0.0f ? vq.y : LEAKY * vq.y);
                pk[2 * q + 1] = pack_f16(z2 > 0.0f ? vq.z : LEAKY * vq.z, z3 > 0.0f ? vq.w : LEAKY * vq.w);
              }
            }
          }
          if (l < 3 || WITH_J) {
            const int c0 = half * 128 + c * 16;
            uint8_t* dstA = sA + (c0 >> 6) * 16384;
            *reinterpret_cast<uint4*>(dstA + sw128_off(row, (c0 & 63) >> 3)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(dstA + sw128_off(row, ((c0 & 63) >> 3) + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
        if (l == 3) part[half * 128 + row] = dot;
        if ((tid & 255) == 0) TL(1, t, l);
        if (l < 3 || WITH_J) signal_a_ready(B, t, rank);
        if ((tid & 255) == 0) TL(2, t, l);
      }
      if (WITH_J) {
        // ---------------- d sdf / d input chain: g_l = (g_{l+1} @ W_{l+1}) * lrelu'(z_l), J = g1 @ W1
#pragma unroll 1
        for (int l = 2; l >= 0; --l) {
          uint32_t bw[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) bw[q] = bits[l * 8 + q];   // issued before the wait: latency hidden
          wait_acc(B, t, acc_par);
          if ((tid & 255) == 0) TL(0, t, 6 - l);
          float v[2][16];
          const int dbg = DBG_MODE;
          if (dbg != 1) tmem_ld16(t_acc, v[0]);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (dbg != 1) {
              tmem_ld_wait();
              if (c < 7) tmem_ld16(t_acc + (c + 1) * 16, v[(c + 1) & 1]);
            }
            float* vv = v[c & 1];
            if (dbg == 2) { if (vv[0] == 123.456f) part[0] = vv[1]; continue; }
            const uint32_t sb = bw[c];
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float m0 = (sb & (2u << (2 * i))) ? LEAKY : 1.0f;
              const float m1 = (sb & (0x20000u << (2 * i))) ? LEAKY : 1.0f;
              pk[i] = pack_f16(vv[2 * i] * m0, vv[2 * i + 1] * m1);
            }
            const int c0 = half * 128 + c * 16;
            uint8_t* dstA = sA + (c0 >> 6) * 16384;
            *reinterpret_cast<uint4*>(dstA + sw128_off(row, (c0 & 63) >> 3)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(dstA + sw128_off(row, ((c0 & 63) >> 3) + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          if ((tid & 255) == 0) TL(1, t, 6 - l);
          signal_a_ready(B, t, rank);
          if ((tid & 255) == 0) TL(2, t, 6 - l);
        }
      } else {
        epi_bar(t);   // `part` visible to the other column half
      }
      // ---------------- Jacobian rows + neighbour interpolation
      float norm = w;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) norm += __shfl_xor_sync(SPF_FULL, norm, o);
      const float wn = sl >= 0 ? w / norm : 0.0f;
      if (WITH_J) {
        wait_acc(B, t, acc_par);
        if ((tid & 255) == 0) TL(0, t, 7);
        if (half == 0) {
          float v[32];
          tmem_ld32(t_row, v);        // d sdf_k / d latent (32 columns)
          tmem_ld_wait();
          if (jw && tile < ntiles) {
            float4* dst = reinterpret_cast<float4*>(jw + ((size_t)tile * 128 + row) * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_float4(wn * v[4 * q], wn * v[4 * q + 1], wn * v[4 * q + 2], wn * v[4 * q + 3]);
          }
          float g[16];
          tmem_ld16(t_row + 32, g);   // columns 32..34 = d sdf_k / d (x - p)
          tmem_ld_wait();
          if (grad) {
            float g0 = wn * g[0], g1 = wn * g[1], g2 = wn * g[2];
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              g0 += __shfl_xor_sync(SPF_FULL, g0, o); g1 += __shfl_xor_sync(SPF_FULL, g1, o); g2 += __shfl_xor_sync(SPF_FULL, g2, o);
            }
            if (sl >= 0 && (row & 7) == 0) { grad[3 * (size_t)sl] = g0; grad[3 * (size_t)sl + 1] = g1; grad[3 * (size_t)sl + 2] = g2; }
          }
        }
      }
      if (half == 0) {
        float agg = wn * (part[row] + part[128 + row] + W.c5);
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) agg += __shfl_xor_sync(SPF_FULL, agg, o);
        if (sl >= 0 && (row & 7) == 0) sdf[sl] = agg;
      }
      tc_fence_before();
      epi_bar(t);   // `part` reads and TMEM reads done before the next iteration reuses them
    }
  }
  teardown(tmem);
}

extern "C" int spf_sdf_fwd_tc(const spf_geo_weights_tc* W, const int32_t* list, const int32_t* count, int64_t n_max,
                              const float* x, const int32_t* pidx, int32_t K, const float* pts, const float* feat_g,
                              float rbf, float* sdf, float* grad, float* jw, void* stream_) {
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


// ------------------------------------------------------------------------------------------------
// colour field forward.  Input columns are permuted (the weight image is packed the same way):
// [c_k (64) | PE6(x - p_k) (39) | 0 (9)], K = 112.  Saved for the backward, by compact pair row: in0 [.,112], h1, h2
// [.,256] (bf16, in the engine's TILE layout -- see signal_a_ready), wn, and the LeakyReLU sign words of z1..z3: m [., 24] (8 words per layer; word w of a layer covers
// columns 32 w .. 32 w + 31: the "negative" bit of column 32 w + 2 i (+1) of the first 16 sits at bit 1 + 2 i (17 + 2 i),
// of the second 16 one position lower).
// ------------------------------------------------------------------------------------------------
#define M_STRIDE 24

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
k_color_fwd_tc2(spf_color_weights_tc W, const int* __restrict__ list, const int* __restrict__ count,
                const float* __restrict__ x, const int* __restrict__ pidx, const float* __restrict__ pts,
                const float* __restrict__ feat_c, float rbf, float* __restrict__ hbar, __nv_bfloat16* __restrict__ in0,
                __nv_bfloat16* __restrict__ h1, __nv_bfloat16* __restrict__ h2, uint32_t* __restrict__ msign,
                float* __restrict__ wn_out, uint8_t* __restrict__ hb_c) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Bars B = carve_bars(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int V = *count;
  const int ntiles = (V + 15) / 16;
  const int nsuper = (ntiles + 3) / 4;
  const int ncl = gridDim.x >> 1, cid = blockIdx.x >> 1;
  const int n_iter = cid < nsuper ? (nsuper - cid + ncl - 1) / ncl : 0;
  Chain& ch = *reinterpret_cast<Chain*>(smem + OFF_CHAIN);
  float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);   // b1..b3
  if (tid == 0) {
    ch.L[0] = {W.w1p, 2, 7, 256, FMT_F16};
    ch.L[1] = {W.w2p, 4, 16, 256, FMT_F16};
    ch.L[2] = {W.w3p, 4, 16, 256, FMT_F16};
    ch.n = 3;
  }
  for (int i = tid; i < 256; i += THREADS) { s_bias[i] = W.b1[i]; s_bias[256 + i] = W.b2[i]; s_bias[512 + i] = W.b3[i]; }
  const uint32_t tmem = setup(smem, B);

  if (warp == WARP_PRODUCER) {
    if (lane == 0) producer_loop(ch, n_iter, rank, smem, B);
  } else if (warp == WARP_MMA) {
    if (rank == 0) mma_loop(ch, n_iter, smem, B, tmem);
    else if (lane == 0) relay_loop(ch, n_iter, B);
  } else {
    const int t = warp >> 3;
    const int row = 32 * (warp & 3) + lane;
    const int half = (warp >> 2) & 1;
    const uint32_t t_acc = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + t * 256 + half * 128;
    uint8_t* sA = smem + OFF_A + t * A_BYTES;
    uint32_t acc_par = 0;
    int sl_n = -1, p_n = -1;
    {
      const int li = (4 * cid + 2 * t + (int)rank) * 16 + (row >> 3);
      if (n_iter > 0 && li < V) { sl_n = list[li]; p_n = pidx[(size_t)sl_n * 8 + (row & 7)]; }
    }
    for (int it = 0; it < n_iter; ++it) {
      const int tile = 4 * (cid + it * ncl) + 2 * t + (int)rank;
      const bool tile_ok = tile < ntiles;
      const size_t grow = (size_t)tile * 128 + row;   // compact pair row
      const int sl = sl_n, p = p_n;
      float xp[3] = {0.f, 0.f, 0.f}, w = 0.f;
      if (p >= 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) xp[a] = x[3 * (size_t)sl + a] - pts[3 * (size_t)p + a];
        w = rbf_w2(xp[0], xp[1], xp[2], rbf);
      }
      float norm = w;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) norm += __shfl_xor_sync(SPF_FULL, norm, o);
      const float wn = sl >= 0 ? w / norm : 0.0f;
      if (half == 0) {
        if (wn_out && tile_ok) wn_out[grow] = wn;
        const float4* src = reinterpret_cast<const float4*>(feat_c + (size_t)(p >= 0 ? p : 0) * 64);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 a = p >= 0 ? src[2 * q] : make_float4(0, 0, 0, 0), b = p >= 0 ? src[2 * q + 1] : make_float4(0, 0, 0, 0);
          uint4 u = make_uint4(pack_f16(a.x, a.y), pack_f16(a.z, a.w), pack_f16(b.x, b.y), pack_f16(b.z, b.w));
          *reinterpret_cast<uint4*>(sA + sw128_off(row, q)) = u;
        }
      } else {
        // PE6 (embedder.py:10-36): [x, sin(2^0 x), cos(2^0 x), ..., sin(2^5 x), cos(2^5 x)] -> 39 values, padded to 48
        float pe[48];
#pragma unroll
        for (int a = 0; a < 3; ++a) pe[a] = xp[a];
        float fr = 1.0f;
#pragma unroll
        for (int l = 0; l < 6; ++l) {
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            float sv, cv;
            sincosf(xp[a] * fr, &sv, &cv);
            pe[3 + 6 * l + a] = p >= 0 ? sv : 0.0f;
            pe[6 + 6 * l + a] = p >= 0 ? cv : 0.0f;
          }
          fr *= 2.0f;
        }
#pragma unroll
        for (int j = 39; j < 48; ++j) pe[j] = 0.0f;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          uint4 u = make_uint4(pack_f16(pe[8 * q], pe[8 * q + 1]), pack_f16(pe[8 * q + 2], pe[8 * q + 3]),
                               pack_f16(pe[8 * q + 4], pe[8 * q + 5]), pack_f16(pe[8 * q + 6], pe[8 * q + 7]));
          *reinterpret_cast<uint4*>(sA + 16384 + sw128_off(row, q)) = u;
        }
      }
      if ((tid & 255) == 0) TL(3, t, it);
      signal_a_ready(B, t, rank, (in0 && tile_ok) ? in0 + (size_t)tile * (2 * 8192) : nullptr, sA, 2 * 16384);
      if ((tid & 255) == 0) TL(4, t, it);
      sl_n = -1; p_n = -1;
      if (it + 1 < n_iter) {
        const int li = (tile + 4 * ncl) * 16 + (row >> 3);
        if (li < V) { sl_n = list[li]; p_n = pidx[(size_t)sl_n * 8 + (row & 7)]; }
      }
#pragma unroll 1
      for (int l = 0; l < 3; ++l) {
        const float4* bias4 = reinterpret_cast<const float4*>(s_bias + l * 256 + half * 128);
        __nv_bfloat16* hdst = (l == 0 ? h1 : h2);
        wait_acc(B, t, acc_par);
        if ((tid & 255) == 0) TL(6, t, l);
        drain_store(t);
        if ((tid & 255) == 0) TL(0, t, l);
        float v[2][16];
        uint32_t sb_prev = 0;
        tmem_ld16(t_acc, v[0]);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          tmem_ld_wait();
          if (c < 7) tmem_ld16(t_acc + (c + 1) * 16, v[(c + 1) & 1]);
          float* vv = v[c & 1];
          const int c0 = half * 128 + c * 16;
          uint32_t pk[8];
          uint32_t sb = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bq = bias4[c * 4 + q];
            float z0 = vv[4 * q] + bq.x, z1 = vv[4 * q + 1] + bq.y, z2 = vv[4 * q + 2] + bq.z, z3 = vv[4 * q + 3] + bq.w;
            // (LeakyReLU on the packed fp16 pair, as in the geometry / head kernels, was measured here and rejected: this
            // kernel also needs the fp32 values in its last layer, and the extra code path cost registers: 0.588 -> 0.632 ms)
            z0 = fmaxf(z0, LEAKY * z0); z1 = fmaxf(z1, LEAKY * z1); z2 = fmaxf(z2, LEAKY * z2); z3 = fmaxf(z3, LEAKY * z3);
            vv[4 * q] = z0; vv[4 * q + 1] = z1; vv[4 * q + 2] = z2; vv[4 * q + 3] = z3;
            pk[2 * q] = pack_f16(z0, z1);
            pk[2 * q + 1] = pack_f16(z2, z3);
            sb = (sb >> 2) | (pk[2 * q] & 0x80008000u);
            sb = (sb >> 2) | (pk[2 * q + 1] & 0x80008000u);
          }
          if (c & 1) {
            if (msign && tile_ok) msign[grow * M_STRIDE + l * 8 + half * 4 + (c >> 1)] = sb_prev | (sb >> 1);
          } else {
            sb_prev = sb;
          }
          if (l < 2) {
            uint8_t* dstA = sA + (c0 >> 6) * 16384;
            const uint4 u0 = make_uint4(pk[0], pk[1], pk[2], pk[3]), u1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            *reinterpret_cast<uint4*>(dstA + sw128_off(row, (c0 & 63) >> 3)) = u0;
            *reinterpret_cast<uint4*>(dstA + sw128_off(row, ((c0 & 63) >> 3) + 1)) = u1;
          } else {
            // hbar[slot][c0..c0+16) = sum over the slot's 8 rows of wn * h3: transpose-reduce over 8 lanes
            float a8[8], a4[4], a2[2];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float lo = wn * vv[j], hi = wn * vv[j + 8];
              const float recv = __shfl_xor_sync(SPF_FULL, (lane & 4) ? lo : hi, 4);
              a8[j] = ((lane & 4) ? hi : lo) + recv;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float recv = __shfl_xor_sync(SPF_FULL, (lane & 2) ? a8[j] : a8[j + 4], 2);
              a4[j] = ((lane & 2) ? a8[j + 4] : a8[j]) + recv;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float recv = __shfl_xor_sync(SPF_FULL, (lane & 1) ? a4[j] : a4[j + 2], 1);
              a2[j] = ((lane & 1) ? a4[j + 2] : a4[j]) + recv;
            }
            if (sl >= 0) {
              const int off = ((lane >> 2) & 1) * 8 + ((lane >> 1) & 1) * 4 + (lane & 1) * 2;
              *reinterpret_cast<float2*>(hbar + (size_t)sl * 256 + c0 + off) = make_float2(a2[0], a2[1]);
              if (hb_c) {   // bf16 copy by COMPACT slot in the tile layout: the radiance head's A operand, loaded by TMA
                const int vs = tile * 16 + (row >> 3), col = c0 + off;
                *reinterpret_cast<uint32_t*>(hb_c + (size_t)(vs >> 7) * (4 * 16384) + (col >> 6) * 16384 +
                                             sw128_off(vs & 127, (col & 63) >> 3) + (col & 7) * 2) = pack_f16(a2[0], a2[1]);
              }
            }
          }
        }
        if ((tid & 255) == 0) TL(1, t, l);
        if (l < 2) signal_a_ready(B, t, rank, (hdst && tile_ok) ? hdst + (size_t)tile * (4 * 8192) : nullptr, sA, 4 * 16384);
        if ((tid & 255) == 0) TL(2, t, l);
      }
      tc_fence_before();
      epi_bar(t);
      if ((tid & 255) == 0) TL(5, t, it);
    }
  }
  teardown(tmem);
}


static int pair_grid(int64_t n_max) {
  const int64_t supers = (n_max + 63) / 64;               // 64 slots = 512 pair rows per cluster iteration
  const int max_cl = spf_num_sms() / 2;
  return 2 * (int)(supers < max_cl ? supers : max_cl);
}

extern "C" int spf_color_fwd_tc(const spf_color_weights_tc* W, const int32_t* list, const int32_t* count, int64_t n_max,
                                const float* x, const int32_t* pidx, int32_t K, const float* pts, const float* feat_c,
                                float rbf, float* hbar, void* in0, void* h1, void* h2, uint32_t* m3, float* wn, void* hb,
                                void* stream_) {
  if (!W || !list || !count || !x || !pidx || !pts || !feat_c || !hbar) return SPF_ERR_INVALID;
  if (K != 8) return SPF_ERR_UNSUPPORTED;
  if (n_max <= 0) return SPF_OK;
  SPF_CUDA(cudaFuncSetAttribute(k_color_fwd_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES), "color_tc2 attr");
  k_color_fwd_tc2<<<pair_grid(n_max), THREADS, SMEM_BYTES, (cudaStream_t)stream_>>>(
      *W, list, count, x, pidx, pts, feat_c, rbf, hbar, (__nv_bfloat16*)in0, (__nv_bfloat16*)h1, (__nv_bfloat16*)h2, m3, wn,
      (uint8_t*)hb);
  SPF_CHECK_LAUNCH("k_color_fwd_tc2");
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
// colour field backward (dgrad chain).  dz rows are written (bf16) for the wgrad GEMMs; d latent is scatter-added.
// ------------------------------------------------------------------------------------------------
// Gradient tiles (dZ) are fp16 scaled by a per-step power of two S (gscale[0]; gscale[1] = 1 / S) chosen on the device
// from the largest upstream gradient so that the chain sits in fp16's normal range (|dZ| ~ 16 at the top: 2^12 of
// headroom for growth through the layers, 2^28 of range below): same tensor throughput as bf16, three more mantissa
// bits, and the same 16-bit format as the saved forward activations -- tcgen05 kind::f16 needs A and B in ONE format, so
// the weight-gradient MMAs consume both as stored.  Everything that leaves the chain (d latent, dW, db) is multiplied by
// 1 / S in fp32; powers of two make that exact.
__device__ __forceinline__ void mask_pack16(const float* vv, uint32_t sb, float scale, uint32_t* pk) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float m0 = (sb & (2u << (2 * i))) ? LEAKY * scale : scale;
    const float m1 = (sb & (0x20000u << (2 * i))) ? LEAKY * scale : scale;
    pk[i] = pack_f16(vv[2 * i] * m0, vv[2 * i + 1] * m1);
  }
}
__device__ __forceinline__ void unpack_f16x8(const uint32_t* w32, float* vv) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(lo), "=f"(hi) : "r"(w32[i]));
    vv[2 * i] = lo;
    vv[2 * i + 1] = hi;
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
k_color_bwd_tc2(spf_color_weights_tc W, const int* __restrict__ list, const int* __restrict__ count,
                const int* __restrict__ pidx, const float* __restrict__ d_hbar, const uint32_t* __restrict__ msign,
                const float* __restrict__ wn_in, __nv_bfloat16* __restrict__ dz1, __nv_bfloat16* __restrict__ dz2,
                __nv_bfloat16* __restrict__ dz3, float* __restrict__ gfeat, const uint8_t* __restrict__ d_hb_c,
                const float* __restrict__ gscale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Bars B = carve_bars(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int V = *count;
  const float gS = gscale[0], gInvS = gscale[1];
  const int ntiles = (V + 15) / 16;
  const int nsuper = (ntiles + 3) / 4;
  const int ncl = gridDim.x >> 1, cid = blockIdx.x >> 1;
  const int n_iter = cid < nsuper ? (nsuper - cid + ncl - 1) / ncl : 0;
  Chain& ch = *reinterpret_cast<Chain*>(smem + OFF_CHAIN);
  if (tid == 0) {
    ch.L[0] = {W.w3tp, 4, 16, 256, FMT_F16};
    ch.L[1] = {W.w2tp, 4, 16, 256, FMT_F16};
    ch.L[2] = {W.w1ftp, 4, 16, 64, FMT_F16};
    ch.n = 3;
  }
  const uint32_t tmem = setup(smem, B);

  if (warp == WARP_PRODUCER) {
    if (lane == 0) producer_loop(ch, n_iter, rank, smem, B);
  } else if (warp == WARP_MMA) {
    if (rank == 0) mma_loop(ch, n_iter, smem, B, tmem);
    else if (lane == 0) relay_loop(ch, n_iter, B);
  } else {
    const int t = warp >> 3;
    const int row = 32 * (warp & 3) + lane;
    const int half = (warp >> 2) & 1;
    const uint32_t t_row = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + t * 256;
    const uint32_t t_acc = t_row + half * 128;
    uint8_t* sA = smem + OFF_A + t * A_BYTES;
    uint32_t acc_par = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int tile = 4 * (cid + it * ncl) + 2 * t + (int)rank;
      const size_t grow = (size_t)tile * 128 + row;
      const int li = tile * 16 + (row >> 3);
      const int sl = li < V ? list[li] : -1;
      const float wn = sl >= 0 ? wn_in[grow] : 0.0f;
      // dz3 = wn * d_hbar[slot] * lrelu'(z3)
      {
        uint4 mw = make_uint4(0, 0, 0, 0);
        if (sl >= 0) mw = *reinterpret_cast<const uint4*>(msign + grow * M_STRIDE + 16 + half * 4);
        const uint32_t mwa[4] = {mw.x, mw.y, mw.z, mw.w};
        const float4* src = reinterpret_cast<const float4*>(d_hbar + (size_t)(sl >= 0 ? sl : 0) * 256 + half * 128);
        // compact mode: row li of the radiance head's bf16 tile-layout gradient (the 8 rows of a slot read the same line)
        const uint8_t* srcc = d_hb_c + (size_t)(li >> 7) * (4 * 16384);
        const int hrow = li & 127;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float vv[16];
          if (d_hb_c) {
            const int c0 = half * 128 + c * 16;
            uint4 a[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
            if (sl >= 0) {
              a[0] = *reinterpret_cast<const uint4*>(srcc + (c0 >> 6) * 16384 + sw128_off(hrow, (c0 & 63) >> 3));
              a[1] = *reinterpret_cast<const uint4*>(srcc + (c0 >> 6) * 16384 + sw128_off(hrow, ((c0 & 63) >> 3) + 1));
            }
            unpack_f16x8(reinterpret_cast<const uint32_t*>(a), vv);   // the head left it scaled by S already
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 d = sl >= 0 ? src[c * 4 + q] : make_float4(0, 0, 0, 0);
              vv[4 * q] = gS * d.x; vv[4 * q + 1] = gS * d.y; vv[4 * q + 2] = gS * d.z; vv[4 * q + 3] = gS * d.w;
            }
          }
          const uint32_t sb = (c & 1) ? (mwa[c >> 1] << 1) : mwa[c >> 1];
          uint32_t pk[8];
          mask_pack16(vv, sb, wn, pk);
          const int c0 = half * 128 + c * 16;
          uint8_t* dstA = sA + (c0 >> 6) * 16384;
          const uint4 u0 = make_uint4(pk[0], pk[1], pk[2], pk[3]), u1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          *reinterpret_cast<uint4*>(dstA + sw128_off(row, (c0 & 63) >> 3)) = u0;
          *reinterpret_cast<uint4*>(dstA + sw128_off(row, ((c0 & 63) >> 3) + 1)) = u1;
        }
      }
      if ((tid & 255) == 0) TL(3, t, it);
      signal_a_ready(B, t, rank, tile < ntiles ? dz3 + (size_t)tile * (4 * 8192) : nullptr, sA, 4 * 16384);
      if ((tid & 255) == 0) TL(4, t, it);
#pragma unroll 1
      for (int l = 0; l < 2; ++l) {
        // dz2 = (dz3 @ W3) * lrelu'(z2) ; dz1 = (dz2 @ W2) * lrelu'(z1)
        uint4 mw = make_uint4(0, 0, 0, 0);
        if (sl >= 0) mw = *reinterpret_cast<const uint4*>(msign + grow * M_STRIDE + (1 - l) * 8 + half * 4);
        const uint32_t mwa[4] = {mw.x, mw.y, mw.z, mw.w};
        __nv_bfloat16* dzo = (l == 0 ? dz2 : dz1);
        wait_acc(B, t, acc_par);
        if ((tid & 255) == 0) TL(6, t, l);
        drain_store(t);
        if ((tid & 255) == 0) TL(0, t, l);
        float v[2][16];
        tmem_ld16(t_acc, v[0]);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          tmem_ld_wait();
          if (c < 7) tmem_ld16(t_acc + (c + 1) * 16, v[(c + 1) & 1]);
          const uint32_t sb = (c & 1) ? (mwa[c >> 1] << 1) : mwa[c >> 1];
          uint32_t pk[8];
          mask_pack16(v[c & 1], sb, 1.0f, pk);
          const int c0 = half * 128 + c * 16;
          uint8_t* dstA = sA + (c0 >> 6) * 16384;
          const uint4 u0 = make_uint4(pk[0], pk[1], pk[2], pk[3]), u1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          *reinterpret_cast<uint4*>(dstA + sw128_off(row, (c0 & 63) >> 3)) = u0;
          *reinterpret_cast<uint4*>(dstA + sw128_off(row, ((c0 & 63) >> 3) + 1)) = u1;
        }
        if ((tid & 255) == 0) TL(1, t, l);
        signal_a_ready(B, t, rank, tile < ntiles ? dzo + (size_t)tile * (4 * 8192) : nullptr, sA, 4 * 16384);
        if ((tid & 255) == 0) TL(2, t, l);
      }
      // d latent = dz1 @ W1[:, latent columns]  -> scatter-add (vector atomics, 16 B each)
      wait_acc(B, t, acc_par);
      if ((tid & 255) == 0) TL(6, t, 2);
      drain_store(t);   // the next iteration's prologue overwrites the A tile
      if ((tid & 255) == 0) TL(0, t, 2);
      {   // both column-half warp sets take 32 of the 64 latent columns each
        const int p = sl >= 0 ? pidx[(size_t)sl * 8 + (row & 7)] : -1;
        float v[32];
        tmem_ld32(t_row + half * 32, v);
        tmem_ld_wait();
        if (p >= 0) {
          float4* dst = reinterpret_cast<float4*>(gfeat + (size_t)p * 64 + half * 32);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            atomicAdd(dst + j4, make_float4(gInvS * v[4 * j4], gInvS * v[4 * j4 + 1], gInvS * v[4 * j4 + 2], gInvS * v[4 * j4 + 3]));
        }
      }
      tc_fence_before();
      epi_bar(t);
      if ((tid & 255) == 0) TL(5, t, it);
    }
  }
  teardown(tmem);
}

extern "C" int spf_color_bwd_tc(const spf_color_weights_tc* W, const int32_t* list, const int32_t* count, int64_t n_max,
                                const int32_t* pidx, int32_t K, const float* d_hbar, const void* h1, const void* h2,
                                const uint32_t* m3, const float* wn, void* dz1, void* dz2, void* dz3, float* feat_c_grad,
                                const void* d_hb, const float* gscale, void* stream_) {
  if (!W || !list || !count || !pidx || (!d_hbar && !d_hb) || !m3 || !wn || !dz1 || !dz2 || !dz3 || !feat_c_grad || !gscale)
    return SPF_ERR_INVALID;
  if (K != 8) return SPF_ERR_UNSUPPORTED;
  if (n_max <= 0) return SPF_OK;
  SPF_CUDA(cudaFuncSetAttribute(k_color_bwd_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES), "colorb_tc2 attr");
  k_color_bwd_tc2<<<pair_grid(n_max), THREADS, SMEM_BYTES, (cudaStream_t)stream_>>>(
      *W, list, count, pidx, d_hbar, m3, wn, (__nv_bfloat16*)dz1, (__nv_bfloat16*)dz2, (__nv_bfloat16*)dz3, feat_c_grad,
      (const uint8_t*)d_hb, gscale);
  SPF_CHECK_LAUNCH("k_color_bwd_tc2");
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
// radiance head on the tile engine (rows are valid SAMPLES: tile = 128 samples, super-tile = 512):
//   f = F_color.6(hbar) ; a1 = lrelu(R.0_f f + zpe[ray]) ; a2 = lrelu(R.2 a1) ; rgb = sigmoid(R.4 a2)
// zpe[ray] = R.0[:, :21] PE3(dir_ray) + R.0.bias is constant per ray and enters as an fp32 per-row bias.
// Saved for the backward by compact sample row: hb (= bf16 hbar), f, a1, a2 [.,256] in the engine's TILE layout (one TMA
// bulk store of the A tile each, see signal_a_ready), pe [.,64] (PE3(dir), 32 columns used, operand of the R.0 wgrad).
// ------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
k_head_fwd_tc2(spf_head_weights_tc W, const int* __restrict__ list, const int* __restrict__ count,
               const float* __restrict__ hbar, const float* __restrict__ zpe, const float* __restrict__ dirs, int Smax,
               float* __restrict__ rgb, __nv_bfloat16* __restrict__ hb_s, __nv_bfloat16* __restrict__ f_s,
               __nv_bfloat16* __restrict__ a1_s, __nv_bfloat16* __restrict__ a2_s, __nv_bfloat16* __restrict__ pe_s) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Bars B = carve_bars(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int V = *count;
  const int ntiles = (V + 127) / 128;
  const int nsuper = (ntiles + 3) / 4;
  const int ncl = gridDim.x >> 1, cid = blockIdx.x >> 1;
  const int n_iter = cid < nsuper ? (nsuper - cid + ncl - 1) / ncl : 0;
  Chain& ch = *reinterpret_cast<Chain*>(smem + OFF_CHAIN);
  float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);   // b4, rb2, rb3
  if (tid == 0) {
    ch.L[0] = {W.w4p, 4, 16, 256, FMT_F16};
    ch.L[1] = {W.r1fp, 4, 16, 256, FMT_F16};
    ch.L[2] = {W.r2p, 4, 16, 256, FMT_F16};
    ch.L[3] = {W.r3p, 4, 16, 32, FMT_F16};
    ch.n = 4;
  }
  for (int i = tid; i < 256; i += THREADS) { s_bias[i] = W.b4[i]; s_bias[256 + i] = W.rb2[i]; }
  if (tid < 3) s_bias[512 + tid] = W.rb3[tid];
  uint64_t* a_load = reinterpret_cast<uint64_t*>(smem + OFF_BAR + 192);   // [2]: A0 of tile t landed (bulk-copy mode)
  if (tid == 0) { mbar_init(a_load, 1); mbar_init(a_load + 1, 1); }       // fenced + synchronised inside setup()
  const uint32_t tmem = setup(smem, B);

  if (warp == WARP_PRODUCER) {
    if (lane == 0) producer_loop(ch, n_iter, rank, smem, B);
  } else if (warp == WARP_MMA) {
    if (rank == 0) mma_loop(ch, n_iter, smem, B, tmem);
    else if (lane == 0) relay_loop(ch, n_iter, B);
  } else {
    const int t = warp >> 3;
    const int row = 32 * (warp & 3) + lane;
    const int half = (warp >> 2) & 1;
    const uint32_t t_row = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + t * 256;
    const uint32_t t_acc = t_row + half * 128;
    uint8_t* sA = smem + OFF_A + t * A_BYTES;
    uint32_t acc_par = 0, ld_par = 0;
    int slot_n = -1;
    {
      const int li = (4 * cid + 2 * t + (int)rank) * 128 + row;
      if (n_iter > 0 && li < V) slot_n = list[li];
    }
    for (int it = 0; it < n_iter; ++it) {
      const int tile = 4 * (cid + it * ncl) + 2 * t + (int)rank;
      const bool tile_ok = tile < ntiles;
      const size_t grow = (size_t)tile * 128 + row;   // compact sample row
      const int slot = slot_n;
      const int ray = slot >= 0 ? slot / Smax : 0;
      if (pe_s && half == 1 && tile_ok) {   // PE3(dir) (embedder.py:10-36): 21 values padded to 32
        float pe[32];
        float d[3] = {0.f, 0.f, 0.f};
        if (slot >= 0) { d[0] = dirs[3 * ray]; d[1] = dirs[3 * ray + 1]; d[2] = dirs[3 * ray + 2]; }
#pragma unroll
        for (int a = 0; a < 3; ++a) pe[a] = d[a];
        float fr = 1.0f;
#pragma unroll
        for (int l = 0; l < 3; ++l) {
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            float sv, cv;
            sincosf(d[a] * fr, &sv, &cv);
            pe[3 + 6 * l + a] = slot >= 0 ? sv : 0.0f;
            pe[6 + 6 * l + a] = slot >= 0 ? cv : 0.0f;
          }
          fr *= 2.0f;
        }
#pragma unroll
        for (int j = 21; j < 32; ++j) pe[j] = 0.0f;
        // tile layout with one k-block (64 columns, 32 used): the R.0 weight gradient reads it like every other operand
        uint8_t* pdst = reinterpret_cast<uint8_t*>(pe_s) + (size_t)tile * 16384;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(pdst + sw128_off(row, q)) =
              make_uint4(pack_f16(pe[8 * q], pe[8 * q + 1]), pack_f16(pe[8 * q + 2], pe[8 * q + 3]),
                         pack_f16(pe[8 * q + 4], pe[8 * q + 5]), pack_f16(pe[8 * q + 6], pe[8 * q + 7]));
      }
      if (hbar) {
        // A0 = hbar[slot] (bf16), gathered by row; saved as hb
        const float4* src = reinterpret_cast<const float4*>(hbar + (size_t)(slot >= 0 ? slot : 0) * 256 + half * 128);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float v[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 dq = slot >= 0 ? src[c * 4 + q] : make_float4(0, 0, 0, 0);
            v[4 * q] = dq.x; v[4 * q + 1] = dq.y; v[4 * q + 2] = dq.z; v[4 * q + 3] = dq.w;
          }
          store_a16(sA, row, half * 128 + c * 16, v);
        }
        if ((tid & 255) == 0) TL(3, t, it);
        signal_a_ready(B, t, rank, (hb_s && tile_ok) ? hb_s + (size_t)tile * (4 * 8192) : nullptr, sA, 4 * 16384);
      } else {
        // A0 = the colour field's compact bf16 copy of hbar, already in the tile layout: one 64 KB bulk copy per tile
        if ((tid & 255) == 0 && tile_ok) {
          mbar_expect_tx(a_load + t, 4 * 16384);
          bulk_g2s(sA, reinterpret_cast<const uint8_t*>(hb_s) + (size_t)tile * (4 * 16384), 4 * 16384, a_load + t);
        }
        bool fix = false;
        if (tile_ok) {
          mbar_wait(a_load + t, ld_par);
          ld_par ^= 1u;
          // rows past the last valid sample of the last tile were never written by the colour kernel: zero them here
          // (and in the global copy, which the F_color.6 weight gradient reads) so that no stale NaN meets a zero dZ
          fix = tile == ntiles - 1 && (V & 127) != 0;
          if (fix && slot < 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c)
              *reinterpret_cast<uint4*>(sA + (half * 2 + (c >> 3)) * 16384 + sw128_off(row, c & 7)) = make_uint4(0u, 0u, 0u, 0u);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c)
            *reinterpret_cast<uint4*>(sA + (half * 2 + (c >> 3)) * 16384 + sw128_off(row, c & 7)) = make_uint4(0u, 0u, 0u, 0u);
        }
        if ((tid & 255) == 0) TL(3, t, it);
        signal_a_ready(B, t, rank, fix ? hb_s + (size_t)tile * (4 * 8192) : nullptr, sA, 4 * 16384);
      }
      if ((tid & 255) == 0) TL(4, t, it);
      slot_n = -1;
      if (it + 1 < n_iter) {
        const int li = (tile + 4 * ncl) * 128 + row;
        if (li < V) {
          slot_n = list[li];
          if (hbar) {   // pull the next tile's hbar half-row (4 lines) towards L2 while this tile's layers run
            const char* nx = reinterpret_cast<const char*>(hbar + (size_t)slot_n * 256 + half * 128);
#pragma unroll
            for (int q = 0; q < 4; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 128 * q));
          }
        }
      }
      const float4* zrow4 = reinterpret_cast<const float4*>(zpe + (size_t)ray * 256 + half * 128);
#pragma unroll 1
      for (int l = 0; l < 3; ++l) {
        const float4* bias4 = l == 1 ? zrow4 : reinterpret_cast<const float4*>(s_bias + (l == 0 ? 0 : 256) + half * 128);
        __nv_bfloat16* dst = l == 0 ? f_s : (l == 1 ? a1_s : a2_s);
        const uint32_t slope2 = l == 0 ? ONE_H2 : LEAKY_H2;   // F_color.6 has no activation
        wait_acc(B, t, acc_par);
        if ((tid & 255) == 0) TL(6, t, l);
        drain_store(t);
        if ((tid & 255) == 0) TL(0, t, l);
        float v[2][16];
        tmem_ld16(t_acc, v[0]);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 bq[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) bq[q] = bias4[c * 4 + q];
          tmem_ld_wait();
          if (c < 7) tmem_ld16(t_acc + (c + 1) * 16, v[(c + 1) & 1]);
          float* vv = v[c & 1];
          uint32_t pk[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float z0 = vv[4 * q] + bq[q].x, z1 = vv[4 * q + 1] + bq[q].y, z2 = vv[4 * q + 2] + bq[q].z, z3 = vv[4 * q + 3] + bq[q].w;
            pk[2 * q] = leaky_f16x2(pack_f16(z0, z1), slope2);
            pk[2 * q + 1] = leaky_f16x2(pack_f16(z2, z3), slope2);
          }
          const int c0 = half * 128 + c * 16;
          uint8_t* dstA = sA + (c0 >> 6) * 16384;
          *reinterpret_cast<uint4*>(dstA + sw128_off(row, (c0 & 63) >> 3)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(dstA + sw128_off(row, ((c0 & 63) >> 3) + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        if ((tid & 255) == 0) TL(1, t, l);
        signal_a_ready(B, t, rank, (dst && tile_ok) ? dst + (size_t)tile * (4 * 8192) : nullptr, sA, 4 * 16384);
        if ((tid & 255) == 0) TL(2, t, l);
      }
      // rgb = sigmoid(R.4 a2 + rb3): 3 of the 32 accumulator columns
      wait_acc(B, t, acc_par);
      if ((tid & 255) == 0) TL(6, t, 3);
      drain_store(t);   // the next iteration's prologue overwrites the A tile
      if (half == 0) {
        float v[16];
        tmem_ld16(t_row, v);
        tmem_ld_wait();
        if (slot >= 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) rgb[3 * (size_t)slot + c] = 1.0f / (1.0f + __expf(-(v[c] + s_bias[512 + c])));
        }
      }
      tc_fence_before();
      epi_bar(t);
      if ((tid & 255) == 0) TL(5, t, it);
    }
  }
  teardown(tmem);
}

static int sample_grid(int64_t n_max) {
  const int64_t supers = (n_max + 511) / 512;             // 512 sample rows per cluster iteration
  const int max_cl = spf_num_sms() / 2;
  return 2 * (int)(supers < max_cl ? supers : max_cl);
}

extern "C" int spf_head_fwd_tc(const spf_head_weights_tc* W, const int32_t* list, const int32_t* count, int64_t n_max,
                               const float* hbar, const float* zpe, const float* dirs, int32_t Smax, float* rgb, void* hb,
                               void* f, void* a1, void* a2, void* pe, void* stream_) {
  if (!W || !list || !count || (!hbar && !hb) || !zpe || !dirs || !rgb || Smax < 1) return SPF_ERR_INVALID;
  if (n_max <= 0) return SPF_OK;
  SPF_CUDA(cudaFuncSetAttribute(k_head_fwd_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES), "head_tc2 attr");
  k_head_fwd_tc2<<<sample_grid(n_max), THREADS, SMEM_BYTES, (cudaStream_t)stream_>>>(
      *W, list, count, hbar, zpe, dirs, Smax, rgb, (__nv_bfloat16*)hb, (__nv_bfloat16*)f, (__nv_bfloat16*)a1,
      (__nv_bfloat16*)a2, (__nv_bfloat16*)pe);
  SPF_CHECK_LAUNCH("k_head_fwd_tc2");
  return SPF_OK;
}

// backward: dz3 = d_rgb * rgb (1 - rgb); dz2 = (dz3 @ R.4) * lrelu'(a2); dz1 = (dz2 @ R.2) * lrelu'(a1);
//           dzf = dz1 @ R.0[:, 21:]; d_hbar = dzf @ F_color.6.   a1 / a2 are read back from the tile layout for their
// signs (sign(a) == sign(z)); dz2, dz1, dzf and dz3 ([rows,64], 16 columns used) leave in the tile layout for the wgrad kernel.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
k_head_bwd_tc2(spf_head_weights_tc W, const int* __restrict__ list, const int* __restrict__ count,
               const float* __restrict__ d_rgb, const float* __restrict__ rgb, const uint8_t* __restrict__ a1_s,
               const uint8_t* __restrict__ a2_s, float* __restrict__ d_hbar, __nv_bfloat16* __restrict__ dzf,
               __nv_bfloat16* __restrict__ dz1, __nv_bfloat16* __restrict__ dz2, __nv_bfloat16* __restrict__ dz3,
               float* __restrict__ drb3, __nv_bfloat16* __restrict__ d_hb_c, const float* __restrict__ gscale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Bars B = carve_bars(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int V = *count;
  const float gS = gscale[0], gInvS = gscale[1];
  const int ntiles = (V + 127) / 128;
  const int nsuper = (ntiles + 3) / 4;
  const int ncl = gridDim.x >> 1, cid = blockIdx.x >> 1;
  const int n_iter = cid < nsuper ? (nsuper - cid + ncl - 1) / ncl : 0;
  Chain& ch = *reinterpret_cast<Chain*>(smem + OFF_CHAIN);
  if (tid == 0) {
    ch.L[0] = {W.r3tp, 1, 1, 256, FMT_F16};
    ch.L[1] = {W.r2tp, 4, 16, 256, FMT_F16};
    ch.L[2] = {W.r1ftp, 4, 16, 256, FMT_F16};
    ch.L[3] = {W.w4tp, 4, 16, 256, FMT_F16};
    ch.n = 4;
  }
  const uint32_t tmem = setup(smem, B);

  if (warp == WARP_PRODUCER) {
    if (lane == 0) producer_loop(ch, n_iter, rank, smem, B);
  } else if (warp == WARP_MMA) {
    if (rank == 0) mma_loop(ch, n_iter, smem, B, tmem);
    else if (lane == 0) relay_loop(ch, n_iter, B);
  } else {
    const int t = warp >> 3;
    const int row = 32 * (warp & 3) + lane;
    const int half = (warp >> 2) & 1;
    const uint32_t t_acc = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + t * 256 + half * 128;
    uint8_t* sA = smem + OFF_A + t * A_BYTES;
    uint32_t acc_par = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int tile = 4 * (cid + it * ncl) + 2 * t + (int)rank;
      const bool tile_ok = tile < ntiles;
      const size_t grow = (size_t)tile * 128 + row;
      const int li = tile * 128 + row;
      const int slot = li < V ? list[li] : -1;
      if (half == 0) {
        float g[3] = {0.f, 0.f, 0.f};
        if (slot >= 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) { const float y = rgb[3 * (size_t)slot + c]; g[c] = gS * d_rgb[3 * (size_t)slot + c] * y * (1.0f - y); }
        }
        const uint4 gz = make_uint4(pack_f16(g[0], g[1]), pack_f16(g[2], 0.f), 0u, 0u);   // dZ of R.4, scaled by S (see mask_pack16)
        if (tile_ok) {   // operand of the R.4 wgrad: tile layout with one k-block (64 columns, 16 used)
          uint8_t* zdst = reinterpret_cast<uint8_t*>(dz3) + (size_t)tile * 16384;
          *reinterpret_cast<uint4*>(zdst + sw128_off(row, 0)) = gz;
          *reinterpret_cast<uint4*>(zdst + sw128_off(row, 1)) = make_uint4(0u, 0u, 0u, 0u);
        }
        if (drb3 && tile_ok) {                                                     // bias gradient of R.4
          const float s0 = warp_sum(g[0]), s1 = warp_sum(g[1]), s2 = warp_sum(g[2]);
          if (lane == 0) { atomicAdd(drb3, gInvS * s0); atomicAdd(drb3 + 1, gInvS * s1); atomicAdd(drb3 + 2, gInvS * s2); }
        }
        *reinterpret_cast<uint4*>(sA + sw128_off(row, 0)) = gz;                    // K = 16: chunks 0, 1 of k-block 0
        *reinterpret_cast<uint4*>(sA + sw128_off(row, 1)) = make_uint4(0u, 0u, 0u, 0u);
      }
      signal_a_ready(B, t, rank);
#pragma unroll 1
      for (int l = 0; l < 3; ++l) {
        const uint8_t* act = (l == 0 ? a2_s : a1_s) + (size_t)tile * (4 * 16384);
        __nv_bfloat16* dzo = l == 0 ? dz2 : (l == 1 ? dz1 : dzf);
        // LeakyReLU masks of this layer, fetched and compressed to one bit per column (bit 16 (c & 1) + j of word c >> 1
        // = "a[c0 + j] > 0") BEFORE waiting for the accumulator: the DRAM latency of the saved activations hides under
        // the MMAs instead of stalling every 16-column chunk of the epilogue
        uint32_t pos[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        if (l < 2 && tile_ok) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int c0 = half * 128 + c * 16;
            const int kb = c0 >> 6, ch0 = (c0 & 63) >> 3;
            const uint4 a0 = *reinterpret_cast<const uint4*>(act + kb * 16384 + sw128_off(row, ch0));
            const uint4 a1 = *reinterpret_cast<const uint4*>(act + kb * 16384 + sw128_off(row, ch0 + 1));
            const uint32_t w32[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            uint32_t b16 = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {   // word i holds activations 2i (low half) and 2i + 1 (high half)
              const uint32_t w = w32[i];
              const uint32_t p0 = (((w & 0x8000u) == 0u) && ((w & 0x7fffu) != 0u)) ? 1u : 0u;
              const uint32_t p1 = (((w & 0x80000000u) == 0u) && ((w & 0x7fff0000u) != 0u)) ? 1u : 0u;
              b16 |= (p0 << (2 * i)) | (p1 << (2 * i + 1));
            }
            if (c & 1) pos[c >> 1] = (pos[c >> 1] & 0x0000ffffu) | (b16 << 16);
            else pos[c >> 1] = (pos[c >> 1] & 0xffff0000u) | b16;
          }
        }
        wait_acc(B, t, acc_par);
        drain_store(t);
        float v[2][16];
        tmem_ld16(t_acc, v[0]);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int c0 = half * 128 + c * 16;
          const int kb = c0 >> 6, ch0 = (c0 & 63) >> 3;
          tmem_ld_wait();
          if (c < 7) tmem_ld16(t_acc + (c + 1) * 16, v[(c + 1) & 1]);
          float* vv = v[c & 1];
          const uint32_t pb = pos[c >> 1] >> (16 * (c & 1));
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float m0 = (pb & (1u << (2 * i))) ? 1.0f : LEAKY;
            const float m1 = (pb & (2u << (2 * i))) ? 1.0f : LEAKY;
            pk[i] = pack_f16(vv[2 * i] * m0, vv[2 * i + 1] * m1);
          }
          uint8_t* dstA = sA + kb * 16384;
          *reinterpret_cast<uint4*>(dstA + sw128_off(row, ch0)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(dstA + sw128_off(row, ch0 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        signal_a_ready(B, t, rank, tile_ok ? dzo + (size_t)tile * (4 * 8192) : nullptr, sA, 4 * 16384);
      }
      // d_hbar = dzf @ F_color.6, consumed by the colour-field backward at valid slots: fp32 by slot, or (d_hb_c) bf16 by
      // COMPACT sample row in the tile layout -- one coalesced 64 KB bulk store per tile instead of 512-byte strided rows
      wait_acc(B, t, acc_par);
      drain_store(t);
      {
        float v[2][16];
        tmem_ld16(t_acc, v[0]);
        float4* dst = reinterpret_cast<float4*>(d_hbar + (size_t)(slot >= 0 ? slot : 0) * 256 + half * 128);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          tmem_ld_wait();
          if (c < 7) tmem_ld16(t_acc + (c + 1) * 16, v[(c + 1) & 1]);
          const float* vv = v[c & 1];
          if (d_hb_c) {
            const int c0 = half * 128 + c * 16;
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = pack_f16(vv[2 * i], vv[2 * i + 1]);   // stays scaled by S for the colour backward
            uint8_t* dstA = sA + (c0 >> 6) * 16384;
            *reinterpret_cast<uint4*>(dstA + sw128_off(row, (c0 & 63) >> 3)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(dstA + sw128_off(row, ((c0 & 63) >> 3) + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          } else if (slot >= 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              dst[c * 4 + q] = make_float4(gInvS * vv[4 * q], gInvS * vv[4 * q + 1], gInvS * vv[4 * q + 2], gInvS * vv[4 * q + 3]);
          }
        }
      }
      tc_fence_before();
      if (d_hb_c) {
        fence_proxy_async();
        epi_bar(t);
        if ((tid & (EPI_THREADS - 1)) == 0) {
          if (tile_ok) { bulk_s2g(d_hb_c + (size_t)tile * (4 * 8192), sA, 4 * 16384); bulk_commit(); }
          bulk_wait_read0();   // the next iteration's prologue overwrites the A tile
        }
      }
      epi_bar(t);
    }
  }
  teardown(tmem);
}

extern "C" int spf_head_bwd_tc(const spf_head_weights_tc* W, const int32_t* list, const int32_t* count, int64_t n_max,
                               const float* d_rgb, const float* rgb, const void* a1, const void* a2, float* d_hbar,
                               void* dzf, void* dz1, void* dz2, void* dz3, float* drb3, void* d_hb, const float* gscale,
                               void* stream_) {
  if (!W || !list || !count || !d_rgb || !rgb || !a1 || !a2 || (!d_hbar && !d_hb) || !dzf || !dz1 || !dz2 || !dz3 || !gscale)
    return SPF_ERR_INVALID;
  if (n_max <= 0) return SPF_OK;
  SPF_CUDA(cudaFuncSetAttribute(k_head_bwd_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES), "headb_tc2 attr");
  k_head_bwd_tc2<<<sample_grid(n_max), THREADS, SMEM_BYTES, (cudaStream_t)stream_>>>(
      *W, list, count, d_rgb, rgb, (const uint8_t*)a1, (const uint8_t*)a2, d_hbar, (__nv_bfloat16*)dzf,
      (__nv_bfloat16*)dz1, (__nv_bfloat16*)dz2, (__nv_bfloat16*)dz3, drb3, (__nv_bfloat16*)d_hb, gscale);
  SPF_CHECK_LAUNCH("k_head_bwd_tc2");
  return SPF_OK;
}

#ifdef SPF_TIMELINE
extern "C" int spf_debug_timeline(unsigned long long* host_out, int max_events) {
  unsigned n = 0;
  cudaMemcpyFromSymbol(&n, eng::g_tl_n, sizeof(n));
  if ((int)n > max_events) n = max_events;
  if (n > 8192) n = 8192;
  cudaMemcpyFromSymbol(host_out, eng::g_tl, sizeof(unsigned long long) * 4 * n);
  unsigned z = 0;
  cudaMemcpyToSymbol(eng::g_tl_n, &z, sizeof(z));
  return (int)n;
}
extern "C" void spf_debug_mode(int m) { cudaMemcpyToSymbol(eng::g_dbg_mode, &m, sizeof(m)); }
#endif
