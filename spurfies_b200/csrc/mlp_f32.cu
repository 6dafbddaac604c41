// mlp_f32.cu -- fp32 ("exact") mode of the per-pair field kernels: gather + RBF weights + MLP chain +
// K-neighbour reduction fused per tile of 8 slots x 8 neighbours = 64 pair rows.
//
// This is the parity anchor (1e-4 vs the reference's fp32 cuBLAS path, which never enables TF32) and the
// numerical reference for the bf16 tcgen05 kernels in mlp_tc.cu.  Plain SIMT FFMA: a 64x256 output tile per
// CTA, 8x8 register micro-tiles, activations ping-pong between two shared-memory buffers, weights streamed
// through a third in 32-row chunks.  Nothing is materialised in HBM except what the backward needs.
//
// Algebraic restructuring w.r.t. the reference graph (same function, fewer FLOPs):
//   * F_geometry.8 and T are both linear with no activation between them (pointneus_disent.py:96-98, 304-306)
//     -> folded on the host into sdf = v5 . h4 + c5;
//   * d sdf / d input is one row vector per pair, so the "autograd.grad(create_graph=True)" pass
//     (pointneus_disent.py:315-323) and the later latent backward are the SAME chain: it is evaluated once in
//     the forward and saved pre-scaled (jw); the backward is a scale + scatter-add;
//   * F_color.6 has no activation (pointneus_disent.py:83) and the RBF weights sum to one, so it commutes
//     with the neighbour interpolation and runs per SAMPLE inside the radiance-head kernel (8x fewer rows).
#include <math.h>
#include "common.cuh"

#define TM 64          // pair rows per tile
#define LD 288         // smem row stride (>= 277 = radiance-head input width, multiple of 32)
#define KC 32          // weight rows per staged chunk
#define NT 256         // threads
#define LEAKY 0.01f
#define SMEM_F32 ((2 * TM * LD + KC * 256) * 4 + 4 * TM * 8 * 4 + 1024)

enum { EPI_BIAS_LEAKY = 0, EPI_BIAS = 1, EPI_MASK_BITS = 2, EPI_MASK_ACT = 3, EPI_NONE = 4 };

// out[TM][0..256) = epi( in[TM][0..kin) @ Wt[kin][ldw] (first n_out columns) )
//   EPI_BIAS_LEAKY : leaky(acc + bias[c]); if bits_out: sign bits of the pre-activation, 8 words per row
//   EPI_BIAS       : acc + bias[c]
//   EPI_MASK_BITS  : acc * (bit(row,c) ? 1 : 0.01)           (bits_in, 8 words per row)
//   EPI_MASK_ACT   : acc * (act[row*256+c] > 0 ? 1 : 0.01)   (saved post-activation values, global)
//   EPI_NONE       : acc
// `in` columns up to round_up(kin, KC) must hold finite values.
template <int EPI>
__device__ __forceinline__ void dense64(const float* __restrict__ in, int kin, const float* __restrict__ Wt, int ldw,
                                        int n_out, const float* __restrict__ bias, float* __restrict__ out,
                                        float* __restrict__ wbuf, uint32_t* bits_out, const uint32_t* bits_in,
                                        const float* __restrict__ act_rows /*global, row 0 of tile*/, int n_rows_valid) {
  const int tid = threadIdx.x, lane = tid & 31, r0 = (tid >> 5) * 8;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  for (int k0 = 0; k0 < kin; k0 += KC) {
    __syncthreads();  // previous chunk consumed / `in` produced
#pragma unroll 8
    for (int kk = 0; kk < KC; ++kk) {
      int k = k0 + kk;
      wbuf[kk * 256 + tid] = (k < kin && tid < n_out) ? Wt[(size_t)k * ldw + tid] : 0.0f;
    }
    __syncthreads();
#pragma unroll 2
    for (int k4 = 0; k4 < KC; k4 += 4) {
      float4 a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(&in[(r0 + i) * LD + k0 + k4]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = wbuf[(k4 + q) * 256 + lane + 32 * j];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float av = q == 0 ? a[i].x : (q == 1 ? a[i].y : (q == 2 ? a[i].z : a[i].w));
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av, w[j], acc[i][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = r0 + i;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = lane + 32 * j;
      float v = acc[i][j];
      if (EPI == EPI_BIAS_LEAKY) {
        v += (c < n_out) ? bias[c] : 0.0f;
        if (bits_out) {
          unsigned b = __ballot_sync(SPF_FULL, v > 0.0f);
          if (lane == 0) bits_out[row * 8 + j] = b;
        }
        v = v > 0.0f ? v : LEAKY * v;
      } else if (EPI == EPI_BIAS) {
        v += (c < n_out) ? bias[c] : 0.0f;
      } else if (EPI == EPI_MASK_BITS) {
        v *= ((bits_in[row * 8 + j] >> lane) & 1u) ? 1.0f : LEAKY;
      } else if (EPI == EPI_MASK_ACT) {
        float a = (row < n_rows_valid) ? act_rows[(size_t)row * 256 + c] : 0.0f;
        v *= a > 0.0f ? 1.0f : LEAKY;
      }
      out[row * LD + c] = v;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float rbf_weight(float dx, float dy, float dz, float rbf) {
  // pointneus_disent.py:241-245 : exp(-(clamp(||x_pi||, 1e-12) * rbf)^2)
  float dist = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
  float tq = dist * rbf;
  return expf(-(tq * tq));
}

// ------------------------------------------------------------------------------------------------
// geometry field: sdf (+ d sdf / d x, + pre-scaled latent Jacobian rows)
// ------------------------------------------------------------------------------------------------
template <bool WITH_J>
__global__ void __launch_bounds__(NT, 1)
k_sdf_f32(spf_geo_weights_f32 W, const int* __restrict__ list, const int* __restrict__ count,
          const float* __restrict__ x, const int* __restrict__ pidx, const float* __restrict__ pts,
          const float* __restrict__ feat_g, float rbf, float* __restrict__ sdf, float* __restrict__ grad,
          float* __restrict__ jw) {
  extern __shared__ __align__(16) float sm[];
  float* bufA = sm;
  float* bufB = bufA + TM * LD;
  float* wbuf = bufB + TM * LD;
  uint32_t* bits = reinterpret_cast<uint32_t*>(wbuf + KC * 256);  // [4][TM][8]
  float* s_w = reinterpret_cast<float*>(bits + 4 * TM * 8);       // [TM]
  float* s_sdf = s_w + TM;                                        // [TM]
  int* s_slot = reinterpret_cast<int*>(s_sdf + TM);               // [8]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int V = *count;
  const int ntiles = (V + 7) / 8;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    if (tid < 8) s_slot[tid] = (tile * 8 + tid < V) ? list[tile * 8 + tid] : -1;
    __syncthreads();
    // ---- gather: 4 threads per row, 8 latent floats each (two float4), part 0 also does x_pi / w
    {
      const int row = tid >> 2, part = tid & 3;
      const int slot = s_slot[row >> 3];
      const int p = slot >= 0 ? pidx[(size_t)slot * 8 + (row & 7)] : -1;
      float4 f0 = make_float4(0, 0, 0, 0), f1 = f0;
      if (p >= 0) {
        const float4* src = reinterpret_cast<const float4*>(feat_g + (size_t)p * 32 + part * 8);
        f0 = src[0]; f1 = src[1];
      }
      float* dst = bufA + row * LD;
      *reinterpret_cast<float4*>(dst + part * 8) = f0;
      *reinterpret_cast<float4*>(dst + part * 8 + 4) = f1;
      if (part == 0) {
        float dx = 0, dy = 0, dz = 0, w = 0;
        if (p >= 0) {
          dx = x[3 * (size_t)slot] - pts[3 * (size_t)p];
          dy = x[3 * (size_t)slot + 1] - pts[3 * (size_t)p + 1];
          dz = x[3 * (size_t)slot + 2] - pts[3 * (size_t)p + 2];
          w = rbf_weight(dx, dy, dz, rbf);
        }
        dst[32] = dx; dst[33] = dy; dst[34] = dz;
        s_w[row] = w;
      }
      for (int c = 35 + part; c < 64; c += 4) dst[c] = 0.0f;
    }
    // ---- forward chain (pointneus_disent.py:86-98, 300-306)
    dense64<EPI_BIAS_LEAKY>(bufA, 35, W.w1t, 256, 256, W.b1, bufB, wbuf, WITH_J ? bits : nullptr, nullptr, nullptr, 0);
    dense64<EPI_BIAS_LEAKY>(bufB, 256, W.w2t, 256, 256, W.b2, bufA, wbuf, WITH_J ? bits + TM * 8 : nullptr, nullptr, nullptr, 0);
    dense64<EPI_BIAS_LEAKY>(bufA, 256, W.w3t, 256, 256, W.b3, bufB, wbuf, WITH_J ? bits + 2 * TM * 8 : nullptr, nullptr, nullptr, 0);
    dense64<EPI_BIAS_LEAKY>(bufB, 256, W.w4t, 256, 256, W.b4, bufA, wbuf, WITH_J ? bits + 3 * TM * 8 : nullptr, nullptr, nullptr, 0);
    // sdf_row = v5 . h4 + c5
    for (int i = 0; i < 8; ++i) {
      const int row = wid * 8 + i;
      float s = 0.0f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s = fmaf(bufA[row * LD + lane + 32 * j], W.v5[lane + 32 * j], s);
      s = warp_sum(s);
      if (lane == 0) s_sdf[row] = s + W.c5;
    }
    __syncthreads();
    if (WITH_J) {
      // ---- d sdf / d input: g4 = v5 * lrelu'(z4); g_l = (g_{l+1} @ W_{l+1}) * lrelu'(z_l); J = g1 @ W1
      for (int i = 0; i < 8; ++i) {
        const int row = wid * 8 + i;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = lane + 32 * j;
          bufB[row * LD + c] = W.v5[c] * (((bits[(3 * TM + row) * 8 + j] >> lane) & 1u) ? 1.0f : LEAKY);
        }
      }
      dense64<EPI_MASK_BITS>(bufB, 256, W.w4, 256, 256, nullptr, bufA, wbuf, nullptr, bits + 2 * TM * 8, nullptr, 0);
      dense64<EPI_MASK_BITS>(bufA, 256, W.w3, 256, 256, nullptr, bufB, wbuf, nullptr, bits + TM * 8, nullptr, 0);
      dense64<EPI_MASK_BITS>(bufB, 256, W.w2, 256, 256, nullptr, bufA, wbuf, nullptr, bits, nullptr, 0);
      dense64<EPI_NONE>(bufA, 256, W.w1, 35, 35, nullptr, bufB, wbuf, nullptr, nullptr, nullptr, 0);
    }
    // ---- neighbour interpolation (pointneus_disent.py:246-247, 308-313): one warp per slot
    {
      const int slot = s_slot[wid];
      if (slot >= 0) {
        float w = lane < 8 ? s_w[wid * 8 + lane] : 0.0f;
        float norm = warp_sum(w);
        float ws = lane < 8 ? w * s_sdf[wid * 8 + lane] : 0.0f;
        float agg = warp_sum(ws);
        if (lane == 0) sdf[slot] = agg / norm;
        if (WITH_J) {
          if (grad && lane < 3) {
            float g = 0.0f;
            for (int k = 0; k < 8; ++k) g += s_w[wid * 8 + k] * bufB[(wid * 8 + k) * LD + 32 + lane];
            grad[3 * (size_t)slot + lane] = g / norm;
          }
          if (jw) {
            for (int k = 0; k < 8; ++k) {
              const int row = wid * 8 + k;
              jw[((size_t)(tile * 8 + wid) * 8 + k) * 32 + lane] = s_w[row] / norm * bufB[row * LD + lane];
            }
          }
        }
      }
    }
  }
}

static int launch_grid(int64_t n_max, int rows_per_tile) {
  int64_t tiles = (n_max + rows_per_tile - 1) / rows_per_tile;
  int sms = spf_num_sms();
  return (int)(tiles < sms ? (tiles > 0 ? tiles : 1) : sms);
}

extern "C" int spf_sdf_fwd_f32(const spf_geo_weights_f32* W, const int32_t* list, const int32_t* count,
                               int64_t n_max, const float* x, const int32_t* pidx, int32_t K, const float* pts,
                               const float* feat_g, float rbf, float* sdf, float* grad, float* jw, void* stream_) {
  if (!W || !list || !count || !x || !pidx || !pts || !feat_g || !sdf) return SPF_ERR_INVALID;
  if (K != 8) return SPF_ERR_UNSUPPORTED;
  if (n_max <= 0) return SPF_OK;
  cudaStream_t st = (cudaStream_t)stream_;
  const bool with_j = grad || jw;
  if (with_j && (!W->w1 || !W->w2 || !W->w3 || !W->w4)) return SPF_ERR_INVALID;
  int grid = launch_grid(n_max, 8);
  if (with_j) {
    SPF_CUDA(cudaFuncSetAttribute(k_sdf_f32<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_F32), "sdf attr");
    k_sdf_f32<true><<<grid, NT, SMEM_F32, st>>>(*W, list, count, x, pidx, pts, feat_g, rbf, sdf, grad, jw);
  } else {
    SPF_CUDA(cudaFuncSetAttribute(k_sdf_f32<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_F32), "sdf attr");
    k_sdf_f32<false><<<grid, NT, SMEM_F32, st>>>(*W, list, count, x, pidx, pts, feat_g, rbf, sdf, grad, jw);
  }
  SPF_CHECK_LAUNCH("k_sdf_f32");
  return SPF_OK;
}

// backward of the geometry field: feat_g_grad[p][:] += d_sdf[slot] * jw[row][:].  Eight lanes per pair row (one float4 of
// the 32 latent channels each, 16-byte vector atomics), four rows per warp and two row groups in flight per iteration: the
// kernel is bound by the dependent list -> pidx -> jw load chain, so bytes in flight per warp are what matters.
__global__ void k_sdf_bwd(const int* __restrict__ list, const int* __restrict__ count, const int* __restrict__ pidx,
                          const float* __restrict__ jw, const float* __restrict__ d_sdf, float* __restrict__ gfeat) {
  const int lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
  const long long nrows = (long long)(*count) * 8;
  const long long wstride = (((long long)gridDim.x * blockDim.x) >> 5) * 8;
  for (long long r0 = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 8; r0 < nrows; r0 += wstride) {
    long long row[2];
    int p[2] = {-1, -1};
    float g[2] = {0.f, 0.f};
    float4 j[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      row[u] = r0 + 4 * u + sub;
      if (row[u] < nrows) {
        const int slot = list[row[u] >> 3];
        p[u] = pidx[(size_t)slot * 8 + (row[u] & 7)];
        g[u] = d_sdf[slot];
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (p[u] >= 0 && g[u] != 0.0f) j[u] = reinterpret_cast<const float4*>(jw)[row[u] * 8 + l8];
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (p[u] >= 0 && g[u] != 0.0f)
        atomicAdd(reinterpret_cast<float4*>(gfeat + (size_t)p[u] * 32) + l8,
                  make_float4(g[u] * j[u].x, g[u] * j[u].y, g[u] * j[u].z, g[u] * j[u].w));
  }
}

extern "C" int spf_sdf_bwd(const int32_t* list, const int32_t* count, int64_t n_max, const int32_t* pidx, int32_t K,
                           const float* jw, const float* d_sdf, float* feat_g_grad, void* stream_) {
  if (!list || !count || !pidx || !jw || !d_sdf || !feat_g_grad) return SPF_ERR_INVALID;
  if (K != 8) return SPF_ERR_UNSUPPORTED;
  if (n_max <= 0) return SPF_OK;
  long long warps = n_max;                 // 8 pair rows (one slot) per warp and iteration
  long long blocks = (warps + 7) / 8;
  long long cap = (long long)spf_num_sms() * 16;
  k_sdf_bwd<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream_>>>(list, count, pidx, jw, d_sdf,
                                                                                        feat_g_grad);
  SPF_CHECK_LAUNCH("k_sdf_bwd");
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
// colour field forward (pointneus_disent.py:325-336), per pair, up to the last activation
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
k_color_fwd_f32(spf_color_weights_f32 W, const int* __restrict__ list, const int* __restrict__ count,
                const float* __restrict__ x, const int* __restrict__ pidx, const float* __restrict__ pts,
                const float* __restrict__ feat_c, float rbf, float* __restrict__ hbar, float* __restrict__ in0,
                float* __restrict__ h1, float* __restrict__ h2, uint32_t* __restrict__ m3, float* __restrict__ wn) {
  extern __shared__ __align__(16) float sm[];
  float* bufA = sm;
  float* bufB = bufA + TM * LD;
  float* wbuf = bufB + TM * LD;
  uint32_t* bits = reinterpret_cast<uint32_t*>(wbuf + KC * 256);  // [TM][8]
  float* s_w = reinterpret_cast<float*>(bits + 4 * TM * 8);
  float* s_wn = s_w + TM;
  int* s_slot = reinterpret_cast<int*>(s_wn + TM);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int V = *count;
  const int ntiles = (V + 7) / 8;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    if (tid < 8) s_slot[tid] = (tile * 8 + tid < V) ? list[tile * 8 + tid] : -1;
    __syncthreads();
    {
      const int row = tid >> 2, part = tid & 3;  // 4 threads per row, 16 colour-latent floats each
      const int slot = s_slot[row >> 3];
      const int p = slot >= 0 ? pidx[(size_t)slot * 8 + (row & 7)] : -1;
      float* dst = bufA + row * LD;
      const float4* src = reinterpret_cast<const float4*>(feat_c + (size_t)(p >= 0 ? p : 0) * 64 + part * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 f = p >= 0 ? src[q] : make_float4(0, 0, 0, 0);
        dst[39 + part * 16 + q * 4 + 0] = f.x; dst[39 + part * 16 + q * 4 + 1] = f.y;
        dst[39 + part * 16 + q * 4 + 2] = f.z; dst[39 + part * 16 + q * 4 + 3] = f.w;
      }
      if (part == 0) {
        float d[3] = {0, 0, 0}, w = 0;
        if (p >= 0) {
#pragma unroll
          for (int a = 0; a < 3; ++a) d[a] = x[3 * (size_t)slot + a] - pts[3 * (size_t)p + a];
          w = rbf_weight(d[0], d[1], d[2], rbf);
        }
        s_w[row] = w;
        // PE6 (embedder.py:10-36): [x, sin(2^0 x), cos(2^0 x), ..., sin(2^5 x), cos(2^5 x)]
#pragma unroll
        for (int a = 0; a < 3; ++a) dst[a] = d[a];
        float fr = 1.0f;
#pragma unroll
        for (int l = 0; l < 6; ++l) {
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            float s, c;
            sincosf(d[a] * fr, &s, &c);
            dst[3 + 6 * l + a] = p >= 0 ? s : 0.0f;
            dst[6 + 6 * l + a] = p >= 0 ? c : 0.0f;
          }
          fr *= 2.0f;
        }
      }
      for (int c = 103 + part; c < 128; c += 4) dst[c] = 0.0f;
    }
    __syncthreads();
    if (tid < 8) {  // normalised weights per slot
      float n = 0.0f;
      for (int k = 0; k < 8; ++k) n += s_w[tid * 8 + k];
      for (int k = 0; k < 8; ++k) s_wn[tid * 8 + k] = s_slot[tid] >= 0 ? s_w[tid * 8 + k] / n : 0.0f;
    }
    __syncthreads();
    const size_t row0 = (size_t)tile * TM;  // compact pair row of this tile's first row
    if (in0) for (int e = tid; e < TM * 104; e += NT) { int r = e / 104, c = e - r * 104; in0[(row0 + r) * 104 + c] = c < 103 ? bufA[r * LD + c] : 0.0f; }
    if (wn && tid < TM) wn[row0 + tid] = s_wn[tid];
    dense64<EPI_BIAS_LEAKY>(bufA, 103, W.w1t, 256, 256, W.b1, bufB, wbuf, nullptr, nullptr, nullptr, 0);
    if (h1) for (int e = tid; e < TM * 256; e += NT) h1[row0 * 256 + e] = bufB[(e >> 8) * LD + (e & 255)];
    dense64<EPI_BIAS_LEAKY>(bufB, 256, W.w2t, 256, 256, W.b2, bufA, wbuf, nullptr, nullptr, nullptr, 0);
    if (h2) for (int e = tid; e < TM * 256; e += NT) h2[row0 * 256 + e] = bufA[(e >> 8) * LD + (e & 255)];
    dense64<EPI_BIAS_LEAKY>(bufA, 256, W.w3t, 256, 256, W.b3, bufB, wbuf, bits, nullptr, nullptr, 0);
    if (m3) for (int e = tid; e < TM * 8; e += NT) m3[row0 * 8 + e] = bits[e];
    // hbar[slot][c] = sum_k wn_k h3[row][c]; thread = column
    for (int s = 0; s < 8; ++s) {
      const int slot = s_slot[s];
      if (slot < 0) continue;
      float a = 0.0f;
#pragma unroll
      for (int k = 0; k < 8; ++k) a = fmaf(s_wn[s * 8 + k], bufB[(s * 8 + k) * LD + tid], a);
      hbar[(size_t)slot * 256 + tid] = a;
    }
  }
}

extern "C" int spf_color_fwd_f32(const spf_color_weights_f32* W, const int32_t* list, const int32_t* count,
                                 int64_t n_max, const float* x, const int32_t* pidx, int32_t K, const float* pts,
                                 const float* feat_c, float rbf, float* hbar, float* in0, float* h1, float* h2,
                                 uint32_t* m3, float* wn, void* stream_) {
  if (!W || !list || !count || !x || !pidx || !pts || !feat_c || !hbar) return SPF_ERR_INVALID;
  if (K != 8) return SPF_ERR_UNSUPPORTED;
  if (n_max <= 0) return SPF_OK;
  SPF_CUDA(cudaFuncSetAttribute(k_color_fwd_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_F32), "color attr");
  k_color_fwd_f32<<<launch_grid(n_max, 8), NT, SMEM_F32, (cudaStream_t)stream_>>>(*W, list, count, x, pidx, pts, feat_c,
                                                                                 rbf, hbar, in0, h1, h2, m3, wn);
  SPF_CHECK_LAUNCH("k_color_fwd_f32");
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
// colour field backward: dgrad chain + latent scatter-add; dz rows are written for the wgrad GEMMs
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
k_color_bwd_f32(spf_color_weights_f32 W, const int* __restrict__ list, const int* __restrict__ count,
                const int* __restrict__ pidx, const float* __restrict__ d_hbar, const float* __restrict__ h1,
                const float* __restrict__ h2, const uint32_t* __restrict__ m3, const float* __restrict__ wn,
                float* __restrict__ dz1, float* __restrict__ dz2, float* __restrict__ dz3,
                float* __restrict__ gfeat) {
  extern __shared__ __align__(16) float sm[];
  float* bufA = sm;
  float* bufB = bufA + TM * LD;
  float* wbuf = bufB + TM * LD;
  int* s_slot = reinterpret_cast<int*>(wbuf + KC * 256);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int V = *count;
  const int ntiles = (V + 7) / 8;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    if (tid < 8) s_slot[tid] = (tile * 8 + tid < V) ? list[tile * 8 + tid] : -1;
    __syncthreads();
    const size_t row0 = (size_t)tile * TM;
    // dz3 = wn * d_hbar[slot] * lrelu'(z3)
    for (int e = tid; e < TM * 256; e += NT) {
      const int r = e >> 8, c = e & 255;
      const int slot = s_slot[r >> 3];
      float v = 0.0f;
      if (slot >= 0) {
        v = wn[row0 + r] * d_hbar[(size_t)slot * 256 + c];
        v *= ((m3[(row0 + r) * 8 + (c >> 5)] >> (c & 31)) & 1u) ? 1.0f : LEAKY;
      }
      bufA[r * LD + c] = v;
      dz3[row0 * 256 + e] = v;
    }
    const int nvalid_rows = min(TM, (V - tile * 8) * 8);
    dense64<EPI_MASK_ACT>(bufA, 256, W.w3, 256, 256, nullptr, bufB, wbuf, nullptr, nullptr, h2 + row0 * 256, nvalid_rows);
    for (int e = tid; e < TM * 256; e += NT) dz2[row0 * 256 + e] = bufB[(e >> 8) * LD + (e & 255)];
    dense64<EPI_MASK_ACT>(bufB, 256, W.w2, 256, 256, nullptr, bufA, wbuf, nullptr, nullptr, h1 + row0 * 256, nvalid_rows);
    for (int e = tid; e < TM * 256; e += NT) dz1[row0 * 256 + e] = bufA[(e >> 8) * LD + (e & 255)];
    // d latent = dz1 @ W1[:, 39:103]
    dense64<EPI_NONE>(bufA, 256, W.w1 + 39, 103, 64, nullptr, bufB, wbuf, nullptr, nullptr, nullptr, 0);
    for (int i = 0; i < 8; ++i) {
      const int r = wid * 8 + i;
      const int slot = s_slot[r >> 3];
      if (slot < 0) continue;
      const int p = pidx[(size_t)slot * 8 + (r & 7)];
      if (p < 0) continue;
      atomicAdd(&gfeat[(size_t)p * 64 + lane], bufB[r * LD + lane]);
      atomicAdd(&gfeat[(size_t)p * 64 + 32 + lane], bufB[r * LD + 32 + lane]);
    }
  }
}

extern "C" int spf_color_bwd_f32(const spf_color_weights_f32* W, const int32_t* list, const int32_t* count,
                                 int64_t n_max, const int32_t* pidx, int32_t K, const float* d_hbar, const float* h1,
                                 const float* h2, const uint32_t* m3, const float* wn, float* dz1, float* dz2,
                                 float* dz3, float* feat_c_grad, void* stream_) {
  if (!W || !list || !count || !pidx || !d_hbar || !h1 || !h2 || !m3 || !wn || !dz1 || !dz2 || !dz3 || !feat_c_grad)
    return SPF_ERR_INVALID;
  if (K != 8) return SPF_ERR_UNSUPPORTED;
  if (n_max <= 0) return SPF_OK;
  SPF_CUDA(cudaFuncSetAttribute(k_color_bwd_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_F32), "colorb attr");
  k_color_bwd_f32<<<launch_grid(n_max, 8), NT, SMEM_F32, (cudaStream_t)stream_>>>(*W, list, count, pidx, d_hbar, h1, h2,
                                                                                 m3, wn, dz1, dz2, dz3, feat_c_grad);
  SPF_CHECK_LAUNCH("k_color_bwd_f32");
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
// radiance head (per sample): f = F_color.6(hbar); rgb = sigmoid(R([PE3(dir) | f]))
// (pointneus_disent.py:83, 100-107, 338-346)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
k_head_fwd_f32(spf_head_weights_f32 W, const int* __restrict__ list, const int* __restrict__ count,
               const float* __restrict__ hbar, const float* __restrict__ dirs, int Smax, float* __restrict__ rgb,
               float* __restrict__ f, float* __restrict__ a1, float* __restrict__ a2) {
  extern __shared__ __align__(16) float sm[];
  float* bufA = sm;
  float* bufB = bufA + TM * LD;
  float* wbuf = bufB + TM * LD;
  int* s_slot = reinterpret_cast<int*>(wbuf + KC * 256);
  const int tid = threadIdx.x;
  const int V = *count;
  const int ntiles = (V + TM - 1) / TM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    if (tid < TM) s_slot[tid] = (tile * TM + tid < V) ? list[tile * TM + tid] : -1;
    __syncthreads();
    const size_t row0 = (size_t)tile * TM;
    for (int e = tid; e < TM * 256; e += NT) {
      const int r = e >> 8, c = e & 255;
      const int slot = s_slot[r];
      bufA[r * LD + c] = slot >= 0 ? hbar[(size_t)slot * 256 + c] : 0.0f;
    }
    dense64<EPI_BIAS>(bufA, 256, W.w4t, 256, 256, W.b4, bufB, wbuf, nullptr, nullptr, nullptr, 0);
    // cat [PE3(dir) (21) | f (256)] into bufA; f is also saved
    for (int e = tid; e < TM * 256; e += NT) {
      const int r = e >> 8, c = e & 255;
      const float v = s_slot[r] >= 0 ? bufB[r * LD + c] : 0.0f;
      bufA[r * LD + 21 + c] = v;
      if (f) f[row0 * 256 + e] = v;
    }
    if (tid < TM) {
      const int slot = s_slot[tid];
      float* dst = bufA + tid * LD;
      float d[3] = {0, 0, 0};
      if (slot >= 0) { const int ray = slot / Smax; d[0] = dirs[3 * ray]; d[1] = dirs[3 * ray + 1]; d[2] = dirs[3 * ray + 2]; }
#pragma unroll
      for (int a = 0; a < 3; ++a) dst[a] = d[a];
      float fr = 1.0f;
#pragma unroll
      for (int l = 0; l < 3; ++l) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          float s, c;
          sincosf(d[a] * fr, &s, &c);
          dst[3 + 6 * l + a] = slot >= 0 ? s : 0.0f;
          dst[6 + 6 * l + a] = slot >= 0 ? c : 0.0f;
        }
        fr *= 2.0f;
      }
      for (int c = 277; c < LD; ++c) dst[c] = 0.0f;
    }
    dense64<EPI_BIAS_LEAKY>(bufA, 277, W.r1t, 256, 256, W.rb1, bufB, wbuf, nullptr, nullptr, nullptr, 0);
    if (a1) for (int e = tid; e < TM * 256; e += NT) a1[row0 * 256 + e] = bufB[(e >> 8) * LD + (e & 255)];
    dense64<EPI_BIAS_LEAKY>(bufB, 256, W.r2t, 256, 256, W.rb2, bufA, wbuf, nullptr, nullptr, nullptr, 0);
    if (a2) for (int e = tid; e < TM * 256; e += NT) a2[row0 * 256 + e] = bufA[(e >> 8) * LD + (e & 255)];
    dense64<EPI_BIAS>(bufA, 256, W.r3t, 3, 3, W.rb3, bufB, wbuf, nullptr, nullptr, nullptr, 0);
    if (tid < TM * 3) {
      const int r = tid / 3, c = tid - 3 * r;
      const int slot = s_slot[r];
      if (slot >= 0) rgb[3 * (size_t)slot + c] = 1.0f / (1.0f + expf(-bufB[r * LD + c]));
    }
  }
}

extern "C" int spf_head_fwd_f32(const spf_head_weights_f32* W, const int32_t* list, const int32_t* count,
                                int64_t n_max, const float* hbar, const float* ray_dirs, int32_t Smax, float* rgb,
                                float* f, float* a1, float* a2, void* stream_) {
  if (!W || !list || !count || !hbar || !ray_dirs || !rgb || Smax < 1) return SPF_ERR_INVALID;
  if (n_max <= 0) return SPF_OK;
  SPF_CUDA(cudaFuncSetAttribute(k_head_fwd_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_F32), "head attr");
  k_head_fwd_f32<<<launch_grid(n_max, TM), NT, SMEM_F32, (cudaStream_t)stream_>>>(*W, list, count, hbar, ray_dirs, Smax,
                                                                                 rgb, f, a1, a2);
  SPF_CHECK_LAUNCH("k_head_fwd_f32");
  return SPF_OK;
}

__global__ void __launch_bounds__(NT, 1)
k_head_bwd_f32(spf_head_weights_f32 W, const int* __restrict__ list, const int* __restrict__ count,
               const float* __restrict__ d_rgb, const float* __restrict__ rgb, const float* __restrict__ a1,
               const float* __restrict__ a2, float* __restrict__ d_hbar, float* __restrict__ dzf,
               float* __restrict__ dz1, float* __restrict__ dz2, float* __restrict__ dz3) {
  extern __shared__ __align__(16) float sm[];
  float* bufA = sm;
  float* bufB = bufA + TM * LD;
  float* wbuf = bufB + TM * LD;
  int* s_slot = reinterpret_cast<int*>(wbuf + KC * 256);
  const int tid = threadIdx.x;
  const int V = *count;
  const int ntiles = (V + TM - 1) / TM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    if (tid < TM) s_slot[tid] = (tile * TM + tid < V) ? list[tile * TM + tid] : -1;
    __syncthreads();
    const size_t row0 = (size_t)tile * TM;
    const int nvalid_rows = min(TM, V - tile * TM);
    // dz3 = d_rgb * rgb (1 - rgb)   (sigmoid')
    for (int e = tid; e < TM * 32; e += NT) {
      const int r = e >> 5, c = e & 31;
      const int slot = s_slot[r];
      float v = 0.0f;
      if (slot >= 0 && c < 3) { float y = rgb[3 * (size_t)slot + c]; v = d_rgb[3 * (size_t)slot + c] * y * (1.0f - y); }
      bufA[r * LD + c] = v;
      if (c < 4) dz3[(row0 + r) * 4 + c] = v;
    }
    dense64<EPI_MASK_ACT>(bufA, 3, W.r3, 256, 256, nullptr, bufB, wbuf, nullptr, nullptr, a2 + row0 * 256, nvalid_rows);
    for (int e = tid; e < TM * 256; e += NT) dz2[row0 * 256 + e] = bufB[(e >> 8) * LD + (e & 255)];
    dense64<EPI_MASK_ACT>(bufB, 256, W.r2, 256, 256, nullptr, bufA, wbuf, nullptr, nullptr, a1 + row0 * 256, nvalid_rows);
    for (int e = tid; e < TM * 256; e += NT) dz1[row0 * 256 + e] = bufA[(e >> 8) * LD + (e & 255)];
    // d f = dz1 @ R1[:, 21:277]
    dense64<EPI_NONE>(bufA, 256, W.r1 + 21, 277, 256, nullptr, bufB, wbuf, nullptr, nullptr, nullptr, 0);
    for (int e = tid; e < TM * 256; e += NT) dzf[row0 * 256 + e] = bufB[(e >> 8) * LD + (e & 255)];
    // d hbar = d f @ W4
    dense64<EPI_NONE>(bufB, 256, W.w4, 256, 256, nullptr, bufA, wbuf, nullptr, nullptr, nullptr, 0);
    for (int e = tid; e < TM * 256; e += NT) {
      const int r = e >> 8, c = e & 255;
      const int slot = s_slot[r];
      if (slot >= 0) d_hbar[(size_t)slot * 256 + c] = bufA[r * LD + c];
    }
  }
}

extern "C" int spf_head_bwd_f32(const spf_head_weights_f32* W, const int32_t* list, const int32_t* count,
                                int64_t n_max, const float* d_rgb, const float* rgb, const float* a1, const float* a2,
                                float* d_hbar, float* dzf, float* dz1, float* dz2, float* dz3, void* stream_) {
  if (!W || !list || !count || !d_rgb || !rgb || !a1 || !a2 || !d_hbar || !dzf || !dz1 || !dz2 || !dz3)
    return SPF_ERR_INVALID;
  if (n_max <= 0) return SPF_OK;
  SPF_CUDA(cudaFuncSetAttribute(k_head_bwd_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_F32), "headb attr");
  k_head_bwd_f32<<<launch_grid(n_max, TM), NT, SMEM_F32, (cudaStream_t)stream_>>>(*W, list, count, d_rgb, rgb, a1, a2,
                                                                                 d_hbar, dzf, dz1, dz2, dz3);
  SPF_CHECK_LAUNCH("k_head_bwd_f32");
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
// exact-mode weight gradient: dW[M][N] += dZ[:, :M]^T @ A[:, :N], db[M] += column sums of dZ, over the first
// count * rows_per_unit compact rows (device-side count: no host synchronisation, CUDA-graph capturable).  Plain fp32
// FFMA, split-K over row partitions with an fp32 atomic merge -- the reference's nn.Linear backward
// (pointneus_disent.py:76-84, 100-107) without a library GEMM.  A's rows may be addressed through an index list
// (row i of dZ pairs with row idx[i] / idx_div of A): the radiance head's hbar (by slot) and PE3(dir) (by ray) operands.
// CTA = 64 x 64 output tile x one row partition; 256 threads, 4 x 4 register micro-tiles, 32-row shared-memory stages.
// ------------------------------------------------------------------------------------------------
#define WGF_ROWS 32
__global__ void __launch_bounds__(256)
k_wgrad_f32(const float* __restrict__ dz, int ldz, int M, const float* __restrict__ act, int lda, int N,
            const int* __restrict__ idx, int idx_div, const int* __restrict__ count, int rows_per_unit, int parts,
            float* __restrict__ dW, float* __restrict__ db) {
  __shared__ __align__(16) float sA[WGF_ROWS][64];
  __shared__ __align__(16) float sB[WGF_ROWS][64];
  __shared__ int sI[WGF_ROWS];
  const long long rows = (long long)(*count) * rows_per_unit;
  const int tiles_n = (N + 63) >> 6;
  const int tm = (int)(blockIdx.y / tiles_n), tn = (int)(blockIdx.y % tiles_n);
  long long chunk = (rows + parts - 1) / parts;
  chunk = (chunk + WGF_ROWS - 1) / WGF_ROWS * WGF_ROWS;
  const long long r0 = (long long)blockIdx.x * chunk, r1 = min(rows, r0 + chunk);
  if (r0 >= r1) return;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  float bsum = 0.0f;
  const int m0 = tm * 64, n0 = tn * 64;
  for (long long r = r0; r < r1; r += WGF_ROWS) {
    const int nr = (int)min((long long)WGF_ROWS, r1 - r);
    __syncthreads();
    if (tid < WGF_ROWS) sI[tid] = (idx && tid < nr) ? idx[r + tid] / idx_div : 0;
    __syncthreads();
    for (int e = tid; e < WGF_ROWS * 64; e += 256) {
      const int k = e >> 6, c = e & 63;
      float a = 0.0f, b = 0.0f;
      if (k < nr) {
        if (m0 + c < M) a = dz[(size_t)(r + k) * ldz + m0 + c];
        if (n0 + c < N) b = act[(size_t)(idx ? (long long)sI[k] : r + k) * lda + n0 + c];
      }
      sA[k][c] = a;
      sB[k][c] = b;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < WGF_ROWS; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sB[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (db && tn == 0 && tid < 64)
      for (int k = 0; k < WGF_ROWS; ++k) bsum += sA[k][tid];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) atomicAdd(dW + (size_t)m * N + n, acc[i][j]);
    }
  if (db && tn == 0 && tid < 64 && m0 + tid < M) atomicAdd(db + m0 + tid, bsum);
}

extern "C" int spf_wgrad_f32(const float* dz, int32_t ldz, int32_t M, const float* act, int32_t lda, int32_t N,
                             const int32_t* idx, int32_t idx_div, const int32_t* count, int32_t rows_per_unit,
                             int64_t n_max, float* dW, float* db, void* stream_) {
  if (!dz || !act || !count || !dW || M < 1 || N < 1 || ldz < M || lda < N || rows_per_unit < 1) return SPF_ERR_INVALID;
  if (idx && idx_div < 1) return SPF_ERR_INVALID;
  if (n_max <= 0) return SPF_OK;
  const int tiles = ((M + 63) / 64) * ((N + 63) / 64);
  const long long max_rows = (long long)n_max * rows_per_unit;
  long long parts = (4LL * spf_num_sms() + tiles - 1) / tiles;           // ~4 CTAs per SM in total
  const long long max_parts = (max_rows + 4 * WGF_ROWS - 1) / (4 * WGF_ROWS);   // at least 128 rows per partition
  if (parts > max_parts) parts = max_parts;
  if (parts < 1) parts = 1;
  dim3 grid((unsigned)parts, (unsigned)tiles);
  k_wgrad_f32<<<grid, 256, 0, (cudaStream_t)stream_>>>(dz, ldz, M, act, lda, N, idx, idx_div, count, rows_per_unit,
                                                        (int)parts, dW, db);
  SPF_CHECK_LAUNCH("k_wgrad_f32");
  return SPF_OK;
}
