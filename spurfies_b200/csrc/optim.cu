// optim.cu -- SURVEY 8(f1): the optimiser step that follows the hot path (spurfies/train.py:355-363, 548-564).
//
//   clip_grad_norm_(params, 1.0)  +  on_after_backward's NaN/Inf guard  +  torch.optim.Adam.step()  +  zero_grad()
//
// over ONE flat fp32 buffer per role (parameters, gradients, exp_avg, exp_avg_sq): two launches, both HBM-bound.
//   k_sumsq_partial / k_sumsq_final : deterministic two-stage sum of squares (fixed grid, fixed order) -> norm_sq[0]
//   k_adam                          : p, g, m, v read once; p, m, v written once; g zeroed for the next step
// The global norm is finite iff every gradient entry is, so the reference's per-tensor isnan/isinf sweep is the same
// reduction.  When it is not finite the whole update is skipped (the reference drops the gradients, and Adam then skips
// every parameter: no moment decay, no step increment).
#include "common.cuh"

#define OPT_THREADS 256
#define OPT_MAX_BLOCKS 1184  // 148 SMs x 8 resident CTAs of 256 threads

__global__ void __launch_bounds__(OPT_THREADS) k_sumsq_partial(const float* __restrict__ g, long long n,
                                                              double* __restrict__ partial) {
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = g4[i];
    float s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    acc += (double)s;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    float v = g[(n4 << 2) + threadIdx.x];
    acc += (double)(v * v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(SPF_FULL, acc, o);
  __shared__ double s_w[OPT_THREADS / 32];
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < OPT_THREADS / 32; ++w) t += s_w[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void k_sumsq_final(const double* __restrict__ partial, int nb, float scale_sq, float* __restrict__ norm_sq) {
  // one warp, fixed order
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb; i += 32) acc += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(SPF_FULL, acc, o);
  if (threadIdx.x == 0) norm_sq[0] = (float)(acc * (double)scale_sq);
}

// state[0] = number of Adam steps taken so far (as float, exact below 2^24), state[1] = learning rate of this step
__global__ void __launch_bounds__(OPT_THREADS) k_adam(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                     float* __restrict__ v, long long n, const float* __restrict__ norm_sq,
                                                     const float* __restrict__ state, float grad_scale, float max_norm,
                                                     double beta1d, double beta2d, float eps, int zero_grad) {
  const float beta1 = (float)beta1d, beta2 = (float)beta2d;
  const float omb1 = (float)(1.0 - beta1d), omb2 = (float)(1.0 - beta2d);
  const float nsq = norm_sq[0];
  const bool finite = isfinite(nsq);
  const float total = sqrtf(nsq);
  // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total + 1e-6), clamped to 1
  float coef = (max_norm > 0.0f) ? fminf(max_norm / (total + 1.0e-6f), 1.0f) : 1.0f;
  coef *= grad_scale;
  const double t = (double)state[0] + 1.0;
  const float lr = state[1];
  // torch/optim/adam.py (_single_tensor_adam / fused): bias corrections in double, applied in fp32
  const float bc1 = (float)(1.0 - pow(beta1d, t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow(beta2d, t));
  const float step_size = lr / bc1;
  const long long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= coef;
    mm = mm + omb1 * (gg - mm);                           // exp_avg.lerp_(grad, 1 - beta1)
    vv = beta2 * vv + omb2 * gg * gg;                     // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp -= (step_size * mm) / denom;
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    if (finite) {
      float4 pv = p4[i], gv = g4[i], mv = m4[i], vv = v4[i];
      upd(pv.x, gv.x, mv.x, vv.x);
      upd(pv.y, gv.y, mv.y, vv.y);
      upd(pv.z, gv.z, mv.z, vv.z);
      upd(pv.w, gv.w, mv.w, vv.w);
      p4[i] = pv; m4[i] = mv; v4[i] = vv;
    }
    if (zero_grad) g4[i] = zero4;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    if (finite) {
      float pv = p[i], mv = m[i], vv = v[i];
      upd(pv, g[i], mv, vv);
      p[i] = pv; m[i] = mv; v[i] = vv;
    }
    if (zero_grad) g[i] = 0.0f;
  }
}

// after k_adam: advance the step counter when the update was applied; info[0] = total norm, info[1] = 1 if skipped
__global__ void k_adam_tick(float* __restrict__ state, const float* __restrict__ norm_sq, float* __restrict__ info) {
  const float nsq = norm_sq[0];
  const bool finite = isfinite(nsq);
  if (finite) state[0] += 1.0f;
  if (info) {
    info[0] = sqrtf(nsq);
    info[1] = finite ? 0.0f : 1.0f;
  }
}

static int opt_blocks(long long n) {
  long long nb = ((n >> 2) + OPT_THREADS - 1) / OPT_THREADS;
  int cap = spf_num_sms() * 8;
  if (cap > OPT_MAX_BLOCKS) cap = OPT_MAX_BLOCKS;
  if (nb < 1) nb = 1;
  return (int)(nb < cap ? nb : cap);
}

extern "C" size_t spf_optim_workspace_bytes(void) { return OPT_MAX_BLOCKS * sizeof(double); }

extern "C" int spf_grad_sumsq(const float* grad, int64_t n, float grad_scale, float* norm_sq, void* workspace,
                              size_t workspace_bytes, void* stream_) {
  if (!grad || !norm_sq || !workspace || n < 0) return SPF_ERR_INVALID;
  if (workspace_bytes < spf_optim_workspace_bytes()) return SPF_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(grad) & 15) != 0) return SPF_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream_;
  const int nb = opt_blocks(n);
  k_sumsq_partial<<<nb, OPT_THREADS, 0, st>>>(grad, (long long)n, (double*)workspace);
  SPF_CHECK_LAUNCH("k_sumsq_partial");
  k_sumsq_final<<<1, 32, 0, st>>>((const double*)workspace, nb, grad_scale * grad_scale, norm_sq);
  SPF_CHECK_LAUNCH("k_sumsq_final");
  return SPF_OK;
}

extern "C" int spf_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                             const float* norm_sq, float* state, float grad_scale, float max_norm, double beta1,
                             double beta2, float eps, int32_t zero_grad, float* info, void* stream_) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !norm_sq || !state || n < 0) return SPF_ERR_INVALID;
  if (((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
        reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) != 0)
    return SPF_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream_;
  k_adam<<<opt_blocks(n), OPT_THREADS, 0, st>>>(param, grad, exp_avg, exp_avg_sq, (long long)n, norm_sq, state, grad_scale,
                                               max_norm, beta1, beta2, eps, zero_grad);
  SPF_CHECK_LAUNCH("k_adam");
  k_adam_tick<<<1, 1, 0, st>>>(state, norm_sq, info);
  SPF_CHECK_LAUNCH("k_adam_tick");
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
// [S, 1 / S] of the tensor-core mode's scaled-fp16 gradient chain (fields.py::grad_scale): S = 2^floor(log2(target /
// max|x|)), clamped to 2^+-100.  One launch instead of the ten elementwise / reduction launches of the torch expression
// (abs, amax, clamp, reciprocal-multiply, log2, floor, clamp, neg, exp2 x2, stack).  |x| as raw bits orders like an
// unsigned integer (NaN above everything, so a NaN gradient gives S = NaN as the torch expression does); the last block
// to finish finalises and leaves the two scratch words zero for the next call.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_grad_scale(const float* __restrict__ x, long long n, float target,
                                                    float* __restrict__ out, unsigned* __restrict__ scratch) {
  unsigned m = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = max(m, __float_as_uint(x[i]) & 0x7fffffffu);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(SPF_FULL, m, o));
  __shared__ unsigned s_m[8];
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = max(m, s_m[w]);
    atomicMax(scratch, m);
    __threadfence();
    if (atomicAdd(scratch + 1, 1u) == gridDim.x - 1) {   // last block
      __threadfence();
      const float raw = __uint_as_float(atomicExch(scratch, 0u));
      const float amax = fmaxf(raw, 1.0e-30f);   // (fmaxf drops a NaN: tested on `raw`)
      scratch[1] = 0u;
      float S, invS;
      if (raw != raw) {
        S = invS = raw;
      } else {
        const float e = fminf(fmaxf(floorf(log2f(target / amax)), -100.0f), 100.0f);
        S = exp2f(e);
        invS = exp2f(-e);
      }
      out[0] = S;
      out[1] = invS;
    }
  }
}

extern "C" int spf_grad_scale(const float* x, int64_t n, float target, float* out, uint32_t* scratch, void* stream_) {
  if (!x || !out || !scratch || n <= 0 || !(target > 0.0f)) return SPF_ERR_INVALID;
  const long long want = (n + 256 * 8 - 1) / (256 * 8);
  const int nb = (int)(want < 296 ? (want < 1 ? 1 : want) : 296);
  k_grad_scale<<<nb, 256, 0, (cudaStream_t)stream_>>>(x, (long long)n, target, out, scratch);
  SPF_CHECK_LAUNCH("k_grad_scale");
  return SPF_OK;
}
