// render.cu -- per-ray kernels: camera rays, filter_points (t / delta), Laplace density + alpha
// compositing (fwd + bwd), VolSDF error-bounded sampler, TV regulariser.
//
// All of these are memory / latency bound segmented scans: ONE WARP PER RAY, lanes strided over the
// ray's samples, prefix sums by warp shuffles, no intermediate tensors (the reference issues ~150
// tiny torch kernels per sampler iteration and a dozen per compositing call).
// Compiled with -fmad=false so that every fp32 operation rounds exactly like the reference's
// separate torch ops (mul, add, div as individual kernels).
#include <math.h>
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// a15: camera rays (spurfies/utils/rend_util.py:60-95, 143-156), pose-matrix branch, B = 1
// ------------------------------------------------------------------------------------------------
__global__ void k_camera_rays(const float* __restrict__ uv, const float* __restrict__ pose,
                              const float* __restrict__ K, int R, float* __restrict__ dirs,
                              float* __restrict__ cam_loc, float* __restrict__ depth_scale) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { cam_loc[0] = pose[3]; cam_loc[1] = pose[7]; cam_loc[2] = pose[11]; }
  if (i >= R) return;
  float fx = K[0], fy = K[5], cx = K[2], cy = K[6], sk = K[1];
  float x = uv[2 * i], y = uv[2 * i + 1];
  // lift (rend_util.py:143-156), z = 1
  float xl = (x - cx + cy * sk / fy - sk * y / fy) / fx * 1.0f;
  float yl = (y - cy) / fy * 1.0f;
  float zl = 1.0f;
  float o[3] = {pose[3], pose[7], pose[11]};
  float w[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) w[a] = pose[4 * a] * xl + pose[4 * a + 1] * yl + pose[4 * a + 2] * zl + o[a];
  float d[3] = {w[0] - o[0], w[1] - o[1], w[2] - o[2]};
  float n = fmaxf(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), 1e-12f);  // F.normalize
  dirs[3 * i] = d[0] / n; dirs[3 * i + 1] = d[1] / n; dirs[3 * i + 2] = d[2] / n;
  // depth_scale = z of the normalised camera-frame direction (pointneus_disent.py:642-645)
  float nc = fmaxf(sqrtf(xl * xl + yl * yl + zl * zl), 1e-12f);
  depth_scale[i] = zl / nc;
}

extern "C" int spf_camera_rays(const float* uv, const float* pose, const float* intrinsics, int32_t R,
                               float* ray_dirs, float* cam_loc, float* depth_scale, void* stream_) {
  if (!uv || !pose || !intrinsics || !ray_dirs || !cam_loc || !depth_scale) return SPF_ERR_INVALID;
  int n = R > 0 ? R : 1;
  k_camera_rays<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(uv, pose, intrinsics, R, ray_dirs, cam_loc,
                                                                     depth_scale);
  SPF_CHECK_LAUNCH("k_camera_rays");
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
// a11a: filter_points (pointneus_disent.py:207-239)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float slot_t(const float* __restrict__ loc, bool valid, const float o[3], const float d[3]) {
  if (!valid) return 0.0f;
  float s = 0.0f;
  int c = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float q = (loc[a] - o[a]) / d[a];  // IEEE division; 0/0 -> NaN dropped by nanmean, x/0 -> inf kept
    if (!isnan(q)) { s += q; c++; }
  }
  return s / (float)c;
}

__global__ void k_ray_prep(const float* __restrict__ sample_loc, const int* __restrict__ pidx,
                           const float* __restrict__ cam_loc, const float* __restrict__ ray_dirs, int R, int Smax,
                           int K, float* __restrict__ t_out, float* __restrict__ delta, float* __restrict__ x_new) {
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= R) return;
  float o[3] = {cam_loc[0], cam_loc[1], cam_loc[2]};
  float d[3] = {ray_dirs[3 * r], ray_dirs[3 * r + 1], ray_dirs[3 * r + 2]};
  for (int s = lane; s < Smax; s += 32) {
    size_t i = (size_t)r * Smax + s;
    bool v = pidx[i * K] >= 0;
    float t = slot_t(sample_loc + 3 * i, v, o, d);
    float tn = 0.0f;
    if (s + 1 < Smax) tn = slot_t(sample_loc + 3 * (i + 1), pidx[(i + 1) * K] >= 0, o, d);
    t_out[i] = t;
    delta[i] = v ? fmaxf(tn - t, 0.0f) : 0.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a) x_new[3 * i + a] = v ? o[a] + t * d[a] : 0.0f;
  }
}

extern "C" int spf_ray_prep(const float* sample_loc, const int32_t* pidx, const float* cam_loc,
                            const float* ray_dirs, int32_t R, int32_t Smax, int32_t K, float* t, float* delta,
                            float* x_new, void* stream_) {
  if (!sample_loc || !pidx || !cam_loc || !ray_dirs || !t || !delta || !x_new) return SPF_ERR_INVALID;
  if (R <= 0) return SPF_OK;
  k_ray_prep<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream_>>>(sample_loc, pidx, cam_loc, ray_dirs, R, Smax, K, t,
                                                             delta, x_new);
  SPF_CHECK_LAUNCH("k_ray_prep");
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
// a10: Laplace density (density.py:21-26) and its partials
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sgnf(float s) { return (s > 0.0f) ? 1.0f : ((s < 0.0f) ? -1.0f : 0.0f); }
__device__ __forceinline__ float laplace_density(float s, float beta) {
  float alpha = 1.0f / beta;
  return alpha * (0.5f + 0.5f * sgnf(s) * expm1f(-fabsf(s) / beta));
}

#define MAX_CHUNKS 4  // Smax <= 128

// ------------------------------------------------------------------------------------------------
// a11b: compositing forward (pointneus_disent.py:894-908, 765-807)
// ------------------------------------------------------------------------------------------------
__global__ void k_composite_fwd(const float* __restrict__ sdf, const float* __restrict__ delta,
                                const float* __restrict__ t, const float* __restrict__ rgb_s,
                                const float* __restrict__ grad, const int* __restrict__ pidx, int K,
                                const int* __restrict__ ray_nvalid, const float* __restrict__ beta_p, int R, int Smax,
                                float* __restrict__ weights, float* __restrict__ rgb, float* __restrict__ depth,
                                float* __restrict__ acc, float* __restrict__ dist, float* __restrict__ normal) {
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= R) return;
  if (ray_nvalid[r] <= 0) {  // pointneus_disent.py:817-835 fill values
    for (int s = lane; s < Smax; s += 32) weights[(size_t)r * Smax + s] = 0.0f;
    if (lane < 3) { rgb[3 * r + lane] = 0.0f; if (normal) normal[3 * r + lane] = 0.0f; }
    if (lane == 0) { depth[r] = 1.0f; acc[r] = 0.0f; dist[r] = 0.0f; }
    return;
  }
  const float beta = beta_p[0];
  float carry = 0.0f;
  float sw = 0.f, swt = 0.f, sc0 = 0.f, sc1 = 0.f, sc2 = 0.f, sn0 = 0.f, sn1 = 0.f, sn2 = 0.f;
  float wloc[MAX_CHUNKS], tloc[MAX_CHUNKS];
  int nch = (Smax + 31) >> 5;
  for (int c = 0; c < nch; ++c) {
    int s = c * 32 + lane;
    size_t i = (size_t)r * Smax + s;
    bool v = s < Smax && pidx[i * K] >= 0;
    float E = 0.0f, ti = 0.0f;
    if (v) { E = delta[i] * laplace_density(sdf[i], beta); ti = t[i]; }
    float inc = warp_scan_incl(E, lane);
    float excl = carry + (inc - E);
    float T = expf(-excl);
    float alpha = 1.0f - expf(-E);
    float w = alpha * T;
    carry += __shfl_sync(SPF_FULL, inc, 31);
    wloc[c] = w; tloc[c] = ti;
    if (s < Smax) weights[i] = w;
    if (v) {
      sw += w; swt += w * ti;
      sc0 += w * rgb_s[3 * i]; sc1 += w * rgb_s[3 * i + 1]; sc2 += w * rgb_s[3 * i + 2];
      if (normal && grad) {
        float gx = grad[3 * i], gy = grad[3 * i + 1], gz = grad[3 * i + 2];
        float gn = sqrtf(gx * gx + gy * gy + gz * gz);  // pointneus_disent.py:805 (no eps)
        sn0 += w * (gx / gn); sn1 += w * (gy / gn); sn2 += w * (gz / gn);
      }
    }
  }
  sw = warp_sum(sw); swt = warp_sum(swt);
  sc0 = warp_sum(sc0); sc1 = warp_sum(sc1); sc2 = warp_sum(sc2);
  // dist_map = sum( w/(sum w + 1e-10) * t )  (pointneus_disent.py:765-770)
  float dm = 0.0f;
  float den = sw + 1e-10f;
  for (int c = 0; c < nch; ++c) dm += wloc[c] / den * tloc[c];
  dm = warp_sum(dm);
  if (normal) { sn0 = warp_sum(sn0); sn1 = warp_sum(sn1); sn2 = warp_sum(sn2); }
  if (lane == 0) {
    rgb[3 * r] = sc0; rgb[3 * r + 1] = sc1; rgb[3 * r + 2] = sc2;
    depth[r] = swt / (sw + 1e-8f);
    acc[r] = sw;
    dist[r] = dm;
    if (normal) { normal[3 * r] = sn0; normal[3 * r + 1] = sn1; normal[3 * r + 2] = sn2; }
  }
}

extern "C" int spf_composite_fwd(const float* sdf, const float* delta, const float* t, const float* rgb_s,
                                 const float* grad, const int32_t* pidx, int32_t K, const int32_t* ray_nvalid,
                                 const float* beta, int32_t R, int32_t Smax, float* weights, float* rgb, float* depth,
                                 float* acc, float* dist, float* normal, void* stream_) {
  if (!sdf || !delta || !t || !rgb_s || !pidx || !ray_nvalid || !beta || !weights || !rgb || !depth || !acc || !dist)
    return SPF_ERR_INVALID;
  if (Smax > 32 * MAX_CHUNKS) return SPF_ERR_UNSUPPORTED;
  if (R <= 0) return SPF_OK;
  k_composite_fwd<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream_>>>(sdf, delta, t, rgb_s, grad, pidx, K, ray_nvalid,
                                                                  beta, R, Smax, weights, rgb, depth, acc, dist,
                                                                  normal);
  SPF_CHECK_LAUNCH("k_composite_fwd");
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
// a11c: compositing + density backward
//   E_i = delta_i sigma_i, T_i = exp(-sum_{j<i} E_j), w_i = (1-e^{-E_i}) T_i
//   dL/dE_k = g_k (T_k - w_k) - sum_{i>k} g_i w_i ,  g_i = total dL/dw_i
// ------------------------------------------------------------------------------------------------
__global__ void k_composite_bwd(const float* __restrict__ sdf, const float* __restrict__ delta,
                                const float* __restrict__ t, const float* __restrict__ rgb_s,
                                const int* __restrict__ pidx, int K, const int* __restrict__ ray_nvalid,
                                const float* __restrict__ beta_p, int R, int Smax, const float* __restrict__ weights,
                                const float* __restrict__ d_weights, const float* __restrict__ d_rgb,
                                const float* __restrict__ d_depth, const float* __restrict__ d_dist,
                                float* __restrict__ d_sdf, float* __restrict__ d_rgb_s, float* __restrict__ d_beta) {
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  float dbeta_acc = 0.0f;
  if (r < R) {
    int nch = (Smax + 31) >> 5;
    if (ray_nvalid[r] <= 0) {
      for (int s = lane; s < Smax; s += 32) {
        size_t i = (size_t)r * Smax + s;
        d_sdf[i] = 0.0f;
        d_rgb_s[3 * i] = 0.f; d_rgb_s[3 * i + 1] = 0.f; d_rgb_s[3 * i + 2] = 0.f;
      }
    } else {
      const float beta = beta_p[0];
      float dr0 = d_rgb ? d_rgb[3 * r] : 0.f, dr1 = d_rgb ? d_rgb[3 * r + 1] : 0.f, dr2 = d_rgb ? d_rgb[3 * r + 2] : 0.f;
      float dd = d_depth ? d_depth[r] : 0.f, dm = d_dist ? d_dist[r] : 0.f;
      float Tl[MAX_CHUNKS], wl[MAX_CHUNKS], gl[MAX_CHUNKS];
      float carry = 0.0f, A = 0.0f, B = 0.0f;
      for (int c = 0; c < nch; ++c) {
        int s = c * 32 + lane;
        size_t i = (size_t)r * Smax + s;
        bool v = s < Smax && pidx[i * K] >= 0;
        float E = v ? delta[i] * laplace_density(sdf[i], beta) : 0.0f;
        float inc = warp_scan_incl(E, lane);
        Tl[c] = expf(-(carry + (inc - E)));
        carry += __shfl_sync(SPF_FULL, inc, 31);
        float w = s < Smax ? weights[i] : 0.0f;
        wl[c] = w;
        if (v) { A += w * t[i]; B += w; }
      }
      A = warp_sum(A); B = warp_sum(B);
      float b8 = B + 1e-8f, b10 = B + 1e-10f;
      float suffix = 0.0f;  // sum of g_i w_i over later chunks
      for (int c = 0; c < nch; ++c) {
        int s = c * 32 + lane;
        size_t i = (size_t)r * Smax + s;
        bool v = s < Smax && pidx[i * K] >= 0;
        float g = 0.0f;
        if (v) {
          float ti = t[i];
          g = (d_weights ? d_weights[i] : 0.0f) + dr0 * rgb_s[3 * i] + dr1 * rgb_s[3 * i + 1] + dr2 * rgb_s[3 * i + 2] +
              dd * (ti / b8 - A / (b8 * b8)) + dm * (ti / b10 - A / (b10 * b10));
        } else if (s < Smax && d_weights) {
          g = d_weights[i];  // invalid slots still carry w = 0, so g never matters; kept for clarity
          g = 0.0f;
        }
        gl[c] = g;
      }
      for (int c = nch - 1; c >= 0; --c) {
        int s = c * 32 + lane;
        size_t i = (size_t)r * Smax + s;
        bool v = s < Smax && pidx[i * K] >= 0;
        float G = gl[c] * wl[c];
        // inclusive suffix scan inside the chunk
        float suf = G;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          float n = __shfl_down_sync(SPF_FULL, suf, o);
          if (lane + o < 32) suf += n;
        }
        float later = suffix + (suf - G);
        suffix += __shfl_sync(SPF_FULL, suf, 0);
        if (s < Smax) {
          float ds = 0.0f;
          if (v) {
            float dE = gl[c] * (Tl[c] - wl[c]) - later;
            float dsig = delta[i] * dE;
            float sv = sdf[i];
            float e = expf(-fabsf(sv) / beta);
            float b2 = beta * beta;
            float dsds, dsdb;
            if (sv > 0.0f) { dsds = -e / (2.0f * b2); dsdb = e * (sv / beta - 1.0f) / (2.0f * b2); }
            else if (sv < 0.0f) { dsds = -e / (2.0f * b2); dsdb = -1.0f / b2 + e * (1.0f + sv / beta) / (2.0f * b2); }
            else { dsds = 0.0f; dsdb = -0.5f / b2; }
            ds = dsig * dsds;
            dbeta_acc += dsig * dsdb;
          }
          d_sdf[i] = ds;
          float w = wl[c];
          d_rgb_s[3 * i] = v ? w * dr0 : 0.f; d_rgb_s[3 * i + 1] = v ? w * dr1 : 0.f; d_rgb_s[3 * i + 2] = v ? w * dr2 : 0.f;
        }
      }
    }
  }
  dbeta_acc = warp_sum(dbeta_acc);
  __shared__ float s_b[8];
  if (lane == 0) s_b[threadIdx.x >> 5] = dbeta_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tsum = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tsum += s_b[w];
    if (tsum != 0.0f) atomicAdd(d_beta, tsum);
  }
}

extern "C" int spf_composite_bwd(const float* sdf, const float* delta, const float* t, const float* rgb_s,
                                 const int32_t* pidx, int32_t K, const int32_t* ray_nvalid, const float* beta,
                                 int32_t R, int32_t Smax, const float* weights, const float* d_weights,
                                 const float* d_rgb, const float* d_depth, const float* d_dist, float* d_sdf,
                                 float* d_rgb_s, float* d_beta, void* stream_) {
  if (!sdf || !delta || !t || !rgb_s || !pidx || !ray_nvalid || !beta || !weights || !d_sdf || !d_rgb_s || !d_beta)
    return SPF_ERR_INVALID;
  if (Smax > 32 * MAX_CHUNKS) return SPF_ERR_UNSUPPORTED;
  if (R <= 0) return SPF_OK;
  k_composite_bwd<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream_>>>(sdf, delta, t, rgb_s, pidx, K, ray_nvalid, beta, R,
                                                                  Smax, weights, d_weights, d_rgb, d_depth, d_dist,
                                                                  d_sdf, d_rgb_s, d_beta);
  SPF_CHECK_LAUNCH("k_composite_bwd");
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
// a12: sampler
// ------------------------------------------------------------------------------------------------
__global__ void k_sampler_coarse(const float* __restrict__ t_vals, const float* __restrict__ t_rand, float near_,
                                 float far_, const float* __restrict__ cam_loc, const float* __restrict__ dirs, int R,
                                 int M, float* __restrict__ z, float* __restrict__ pts) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)R * M) return;
  int r = (int)(i / M), m = (int)(i - (long long)r * M);
  // ray_sampler.py:46-47
  float tv = t_vals[m];
  float zc = near_ * (1.0f - tv) + far_ * tv;
  float zv = zc;
  if (t_rand) {  // ray_sampler.py:49-57 stratified
    float lower = zc, upper = zc;
    if (m > 0) { float tp = t_vals[m - 1]; float zp = near_ * (1.0f - tp) + far_ * tp; lower = 0.5f * (zc + zp); }
    if (m < M - 1) { float tn = t_vals[m + 1]; float zn = near_ * (1.0f - tn) + far_ * tn; upper = 0.5f * (zn + zc); }
    zv = lower + (upper - lower) * t_rand[i];
  }
  z[i] = zv;
#pragma unroll
  for (int a = 0; a < 3; ++a) pts[3 * i + a] = cam_loc[a] + zv * dirs[3 * r + a];  // ray_sampler.py:398
}

extern "C" int spf_sampler_coarse(const float* t_vals, const float* t_rand, float near_, float far_,
                                  const float* cam_loc, const float* ray_dirs, int32_t R, int32_t M, float* z,
                                  float* points, void* stream_) {
  if (!t_vals || !cam_loc || !ray_dirs || !z || !points || M < 2) return SPF_ERR_INVALID;
  if (R <= 0) return SPF_OK;
  long long n = (long long)R * M;
  k_sampler_coarse<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(t_vals, t_rand, near_, far_,
                                                                                  cam_loc, ray_dirs, R, M, z, points);
  SPF_CHECK_LAUNCH("k_sampler_coarse");
  return SPF_OK;
}

// error bound for one beta (ray_sampler.py:576-588); lanes own contiguous chunks of the M-1 sections.  The per-section
// terms (error-integral increment, free-energy increment) are needed twice -- for the chunk sums that feed the warp scans
// and again when the lane walks its chunk -- so they are parked in the lane's own slice of two shared scratch rows
// (se_s, sf_s) instead of being recomputed (two exps, an expm1 and three divisions per section and call).
__device__ float error_bound_warp(const float* __restrict__ sz, const float* __restrict__ sd,
                                  const float* __restrict__ sds, float* __restrict__ se_s, float* __restrict__ sf_s,
                                  int M, float beta, int lane) {
  const int n = M - 1;
  const int ch = (n + 31) >> 5;
  const int i0 = lane * ch, i1 = min(n, i0 + ch);
  const float fb2 = 4.0f * (beta * beta);
  float se = 0.0f, sf = 0.0f;
  for (int i = i0; i < i1; ++i) {
    float a = sz[i + 1] - sz[i];
    const float ei = expf(-sds[i] / beta) * (a * a) / fb2;
    const float fi = a * laplace_density(sd[i], beta);
    se_s[i] = ei; sf_s[i] = fi;
    se += ei;
    sf += fi;
  }
  float ie = warp_scan_incl(se, lane) - se;  // exclusive offsets
  float jf = warp_scan_incl(sf, lane) - sf;
  float mx = -INFINITY;
  bool anynan = false;
  for (int i = i0; i < i1; ++i) {
    ie += se_s[i];                                                // inclusive error integral
    float b = (fminf(expf(ie), 1.0e6f) - 1.0f) * expf(-jf);       // uses the exclusive density integral
    jf += sf_s[i];
    anynan |= isnan(b);
    mx = fmaxf(mx, b);
  }
  mx = warp_max(mx);
  if (__any_sync(SPF_FULL, anynan)) mx = NAN;  // torch.max propagates NaN
  return mx;
}

__global__ void k_sampler_iter(const float* __restrict__ z, const float* __restrict__ sdf, int R, int M,
                               const float* __restrict__ beta_dev, float eps, int beta_iters, float bound_coef,
                               float add_tiny, int first_iter, float* __restrict__ beta_io, int final_, int N,
                               const float* __restrict__ u, const float* __restrict__ u_lin, float near_, float far_,
                               const int* __restrict__ extra_idx, int n_extra, const float* __restrict__ cam_loc,
                               const float* __restrict__ dirs, float* __restrict__ out_z, float* __restrict__ out_pts,
                               int* __restrict__ flag, const int* __restrict__ state, int pred) {
  extern __shared__ float smem[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + wid;
  if (pred && r < R) {
    // device-side control flow of Algorithm 1's outer loop (ray_sampler.py:466-474) for the multi-iteration (eval)
    // schedule: state[0] = "some ray's beta is still above beta0" (set by this iteration's probe launch), state[1] =
    // "converged in an earlier iteration, the final draw has been made".  No host round trip decides what runs.
    const int not_conv = state[0], done = state[1];
    if (pred == 1 && done) {   // probe after convergence: park the new samples outside every grid (the SDF pass on
      for (int j = lane; j < N; j += 32) {   // them then finds no neighbour and does no MLP work)
        out_z[(size_t)r * N + j] = far_;
#pragma unroll
        for (int a = 0; a < 3; ++a) out_pts[((size_t)r * N + j) * 3 + a] = 1.0e9f;
      }
      return;
    }
    if (pred == 2 && (done || not_conv)) return;   // final draw only in the iteration that converged
    if (pred == 3 && done) return;                 // last allowed iteration: final draw unless already made
  }
  const int stride = M + 1;
  float* sz = smem + (size_t)wid * 6 * stride;   // z
  float* sd = sz + stride;                        // sdf
  float* sds = sd + stride;                       // d_star, later scratch for the final sort
  float* scdf = sds + stride;                     // cdf
  float* se_s = scdf + stride;                    // per-section error-integral increments (lane-private slices)
  float* sf_s = se_s + stride;                    // per-section / per-sample free-energy increments
  if (r >= R) return;
  for (int i = lane; i < M; i += 32) { sz[i] = z[(size_t)r * M + i]; sd[i] = sdf[(size_t)r * M + i]; }
  __syncwarp();
  const int n = M - 1;
  // d* (ray_sampler.py:417-432) and sum of squared dists
  float ss = 0.0f;
  for (int i = lane; i < n; i += 32) {
    float a = sz[i + 1] - sz[i];
    float d0 = sd[i], d1 = sd[i + 1];
    float b = fabsf(d0), c = fabsf(d1);
    bool first = a * a + b * b <= c * c;
    bool second = a * a + c * c <= b * b;
    float ds = 0.0f;
    if (first) ds = b;
    if (second) ds = c;
    float s = (a + b + c) / 2.0f;
    float area = s * (s - a) * (s - b) * (s - c);
    if (!first && !second && (b + c - a > 0.0f)) ds = (2.0f * sqrtf(area)) / a;
    if (!(sgnf(d1) * sgnf(d0) == 1.0f)) ds = 0.0f;
    sds[i] = ds;
    ss += a * a;
  }
  ss = warp_sum(ss);
  __syncwarp();
  const float beta0 = beta_dev[0];
  float beta = first_iter ? sqrtf(bound_coef * ss) : beta_io[r];  // ray_sampler.py:388-392
  // line search (ray_sampler.py:435-445)
  float e0 = error_bound_warp(sz, sd, sds, se_s, sf_s, M, beta0, lane);
  if (e0 <= eps) beta = beta0;
  float bmin = beta0, bmax = beta;
  // A ray that meets the bound at beta0 (every ray that misses the cloud, and every converged ray of a later iteration)
  // starts the bisection on the empty interval [beta0, beta0]: each of its steps re-evaluates the bound at mid = beta0,
  // finds e0 again and leaves bmax = beta0.  Skipping them is bit-identical and saves 10 of the ray's 11 evaluations
  // (e0 is warp-uniform: a warp owns the ray).
  if (!(e0 <= eps)) {
    for (int j = 0; j < beta_iters; ++j) {
      float mid = (bmin + bmax) / 2.0f;
      float e = error_bound_warp(sz, sd, sds, se_s, sf_s, M, mid, lane);
      if (e <= eps) bmax = mid;
      else if (e > eps) bmin = mid;
    }
  }
  beta = bmax;
  if (lane == 0) {
    beta_io[r] = beta;
    if (beta > beta0) atomicOr(flag, 1);  // ray_sampler.py:468
  }
  // weights / transmittance with the chosen beta (ray_sampler.py:448-464), lanes own contiguous chunks of M
  const int chM = (M + 31) >> 5;
  const int j0 = lane * chM, j1 = min(M, j0 + chM);
  float sf = 0.0f, se = 0.0f;
  const float fb2 = 4.0f * (beta * beta);
  __syncwarp();   // the chunking below differs from error_bound_warp's: its scratch reads must be done
  for (int i = j0; i < j1; ++i) {
    float a = i < n ? sz[i + 1] - sz[i] : 1e10f;
    const float fi = a * laplace_density(sd[i], beta);
    sf_s[i] = fi;
    sf += fi;
    if (i < n) { const float ei = expf(-sds[i] / beta) * (a * a) / fb2; se_s[i] = ei; se += ei; }
  }
  float jf = warp_scan_incl(sf, lane) - sf;
  float ie = warp_scan_incl(se, lane) - se;
  // unnormalised pdf into scdf[0..n)
  float tot = 0.0f;
  for (int i = j0; i < j1; ++i) {
    float fe = sf_s[i];
    float T = expf(-jf);
    float p;
    if (final_) {
      float w = (1.0f - expf(-fe)) * T;
      p = w + 1e-5f;                                               // ray_sampler.py:495-497
    } else {
      if (i < n) ie += se_s[i];
      p = (fminf(expf(ie), 1.0e6f) - 1.0f) * T + add_tiny;         // ray_sampler.py:476-486
    }
    jf += fe;
    if (i < n) { scdf[i + 1] = p; tot += p; }
  }
  tot = warp_sum(tot);
  __syncwarp();
  // normalise + inclusive scan -> cdf[0..M)  (cdf[0] = 0)
  {
    const int ch = (n + 31) >> 5;
    const int i0 = lane * ch, i1 = min(n, i0 + ch);
    float s = 0.0f;
    for (int i = i0; i < i1; ++i) { float p = scdf[i + 1] / tot; scdf[i + 1] = p; s += p; }
    float off = warp_scan_incl(s, lane) - s;
    for (int i = i0; i < i1; ++i) { off += scdf[i + 1]; scdf[i + 1] = off; }
    if (lane == 0) scdf[0] = 0.0f;
  }
  __syncwarp();
  // inverse CDF (ray_sampler.py:517-529)
  float* ssort = sds;  // reuse
  const int cols = final_ ? N + 2 + n_extra : N;
  for (int j = lane; j < N; j += 32) {
    float uu = u ? u[(size_t)r * N + j] : u_lin[j];
    int lo = 0, hi = M;  // first index with cdf > uu  (searchsorted right=True)
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (scdf[mid] <= uu) lo = mid + 1; else hi = mid;
    }
    int below = max(lo - 1, 0), above = min(M - 1, lo);
    float cb = scdf[below], ca = scdf[above];
    float den = ca - cb;
    if (den < 1e-5f) den = 1.0f;
    float tt = (uu - cb) / den;
    float smp = sz[below] + tt * (sz[above] - sz[below]);
    if (final_) ssort[j] = smp;
    else {
      out_z[(size_t)r * N + j] = smp;
#pragma unroll
      for (int a = 0; a < 3; ++a) out_pts[((size_t)r * N + j) * 3 + a] = cam_loc[a] + smp * dirs[3 * r + a];
    }
  }
  if (!final_) return;
  // extras + sort (ray_sampler.py:548-559)
  if (lane == 0) { ssort[N] = near_; ssort[N + 1] = far_; }
  for (int e = lane; e < n_extra; e += 32) ssort[N + 2 + e] = sz[extra_idx[e]];
  __syncwarp();
  bool mynan = false;
  for (int i = lane; i < cols; i += 32) mynan |= ssort[i] != ssort[i];
  const bool row_has_nan = __any_sync(SPF_FULL, mynan);
  for (int i = lane; i < cols; i += 32) {
    float v = ssort[i];
    int rank = 0;
    if (!row_has_nan) {   // the common case: plain stable rank (ties by position)
      for (int j = 0; j < cols; ++j) {
        float o = ssort[j];
        rank += (o < v) || (o == v && j < i);
      }
    } else {
      const bool vnan = v != v;
      for (int j = 0; j < cols; ++j) {
        // torch.sort order: NaN sorts last (a miss ray's samples are NaN, as in the reference); ties by position, so
        // the ranks are a permutation and every output element is written
        float o = ssort[j];
        const bool onan = o != o;
        rank += (!onan && (vnan || o < v)) || ((o == v || (onan && vnan)) && j < i);
      }
    }
    out_z[(size_t)r * cols + rank] = v;
#pragma unroll
    for (int a = 0; a < 3; ++a) out_pts[((size_t)r * cols + rank) * 3 + a] = cam_loc[a] + v * dirs[3 * r + a];
  }
}

extern "C" int spf_sampler_iter(const float* z, const float* sdf, int32_t R, int32_t M, const float* beta_dev,
                                float eps, int32_t beta_iters, float bound_coef, float add_tiny, int32_t first_iter,
                                float* beta_io, int32_t final_, int32_t N, const float* u, const float* u_lin,
                                float near_, float far_, const int32_t* extra_idx, int32_t n_extra,
                                const float* cam_loc, const float* ray_dirs, float* out_z, float* out_points,
                                int32_t* flag_not_converged, void* stream_) {
  if (!z || !sdf || !beta_dev || !beta_io || !cam_loc || !ray_dirs || !out_z || !out_points || !flag_not_converged)
    return SPF_ERR_INVALID;
  if (!u && !u_lin) return SPF_ERR_INVALID;
  if (M < 2 || N < 1) return SPF_ERR_INVALID;
  if (final_ && (N + 2 + n_extra > M)) return SPF_ERR_UNSUPPORTED;
  if (n_extra > 0 && !extra_idx) return SPF_ERR_INVALID;
  if (R <= 0) return SPF_OK;
  const int wpb = 4;
  size_t smem = (size_t)wpb * 6 * (M + 1) * sizeof(float);
  if (smem > 200 * 1024) return SPF_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    SPF_CUDA(cudaFuncSetAttribute(k_sampler_iter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
             "sampler_iter smem attr");
  k_sampler_iter<<<(R + wpb - 1) / wpb, wpb * 32, smem, (cudaStream_t)stream_>>>(
      z, sdf, R, M, beta_dev, eps, beta_iters, bound_coef, add_tiny, first_iter, beta_io, final_, N, u, u_lin, near_,
      far_, extra_idx, n_extra, cam_loc, ray_dirs, out_z, out_points, flag_not_converged, nullptr, 0);
  SPF_CHECK_LAUNCH("k_sampler_iter");
  return SPF_OK;
}

extern "C" int spf_sampler_iter_pred(const float* z, const float* sdf, int32_t R, int32_t M, const float* beta_dev,
                                     float eps, int32_t beta_iters, float bound_coef, float add_tiny, int32_t first_iter,
                                     float* beta_io, int32_t final_, int32_t N, const float* u, const float* u_lin,
                                     float near_, float far_, const int32_t* extra_idx, int32_t n_extra,
                                     const float* cam_loc, const float* ray_dirs, float* out_z, float* out_points,
                                     int32_t* state, int32_t pred, void* stream_) {
  if (!z || !sdf || !beta_dev || !beta_io || !cam_loc || !ray_dirs || !out_z || !out_points || !state) return SPF_ERR_INVALID;
  if (!u && !u_lin) return SPF_ERR_INVALID;
  if (M < 2 || N < 1 || pred < 1 || pred > 3) return SPF_ERR_INVALID;
  if ((pred == 1) != (final_ == 0)) return SPF_ERR_INVALID;
  if (final_ && (N + 2 + n_extra > M)) return SPF_ERR_UNSUPPORTED;
  if (n_extra > 0 && !extra_idx) return SPF_ERR_INVALID;
  if (R <= 0) return SPF_OK;
  const int wpb = 4;
  size_t smem = (size_t)wpb * 6 * (M + 1) * sizeof(float);
  if (smem > 200 * 1024) return SPF_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    SPF_CUDA(cudaFuncSetAttribute(k_sampler_iter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
             "sampler_iter smem attr");
  k_sampler_iter<<<(R + wpb - 1) / wpb, wpb * 32, smem, (cudaStream_t)stream_>>>(
      z, sdf, R, M, beta_dev, eps, beta_iters, bound_coef, add_tiny, first_iter, beta_io, final_, N, u, u_lin, near_,
      far_, extra_idx, n_extra, cam_loc, ray_dirs, out_z, out_points, state, state, pred);
  SPF_CHECK_LAUNCH("k_sampler_iter");
  return SPF_OK;
}

// merge two sorted rows (ray_sampler.py:533 sort of cat, :405-415 gather of the merged sdf)
__global__ void k_sampler_merge(const float* __restrict__ z, const float* __restrict__ sdf, int M,
                                const float* __restrict__ zs, const float* __restrict__ sdf_s, int N, int R,
                                float* __restrict__ z_out, float* __restrict__ sdf_out) {
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* a = z + (size_t)r * M;
  const float* b = zs + (size_t)r * N;
  float* zo = z_out + (size_t)r * (M + N);
  float* so = sdf_out + (size_t)r * (M + N);
  // torch.sort order with NaN last (the samples of a miss ray are NaN): the two rank maps stay a permutation
  for (int i = lane; i < M; i += 32) {
    float v = a[i];
    const bool vnan = v != v;
    int lo = 0, hi = N;  // count of b sorting strictly before v
    while (lo < hi) { int mid = (lo + hi) >> 1; float x = b[mid]; if ((x == x) && (vnan || x < v)) lo = mid + 1; else hi = mid; }
    zo[i + lo] = v; so[i + lo] = sdf[(size_t)r * M + i];
  }
  for (int j = lane; j < N; j += 32) {
    float v = b[j];
    const bool vnan = v != v;
    int lo = 0, hi = M;  // count of a sorting before or equal to v
    while (lo < hi) { int mid = (lo + hi) >> 1; if (vnan || a[mid] <= v) lo = mid + 1; else hi = mid; }
    zo[j + lo] = v; so[j + lo] = sdf_s[(size_t)r * N + j];
  }
}

extern "C" int spf_sampler_merge(const float* z, const float* sdf, int32_t M, const float* zs, const float* sdf_s,
                                 int32_t N, int32_t R, float* z_out, float* sdf_out, void* stream_) {
  if (!z || !sdf || !zs || !sdf_s || !z_out || !sdf_out) return SPF_ERR_INVALID;
  if (R <= 0) return SPF_OK;
  k_sampler_merge<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream_>>>(z, sdf, M, zs, sdf_s, N, R, z_out, sdf_out);
  SPF_CHECK_LAUNCH("k_sampler_merge");
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
// a13: tv_regul (spurfies/model/utils.py:221-281) over cached self-kNN lists; one warp per point,
// one lane per latent channel (C_g = 32).
// ------------------------------------------------------------------------------------------------
__global__ void k_tv(const float* __restrict__ pts, const float* __restrict__ feat, const int* __restrict__ nbr, int N,
                     int K, int first, int count, float* __restrict__ value, float* __restrict__ grad, float grad_scale) {
  // points [first, first + count) of the N: a data-parallel rank takes one slice of the (ray-independent) regulariser
  const int li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int i = first + li;
  int lane = threadIdx.x & 31;
  float tv_i = 0.0f;
  const bool active = li < count && i < N;
  if (active) {
    // utils.py:236-253: pad with self, drop self when other neighbours exist
    int my = lane < K ? nbr[(size_t)i * K + lane] : -1;
    bool has = __any_sync(SPF_FULL, my >= 0);
    if (!has && lane == 0) my = i;
    int cnt = __popc(__ballot_sync(SPF_FULL, my >= 0));
    if (cnt > 1 && my == i) my = -1;
    float w = 0.0f;
    if (my >= 0) {
      float dx = pts[3 * my] - pts[3 * i], dy = pts[3 * my + 1] - pts[3 * i + 1], dz = pts[3 * my + 2] - pts[3 * i + 2];
      w = 1.0f / (sqrtf(dx * dx + dy * dy + dz * dz) + 1.0e-5f);
    }
    float norm = warp_sum(w);
    float fi = feat[(size_t)i * 32 + lane];
    float gi = 0.0f;
    float part = 0.0f;   // this channel's share of sum_k w_k |f_j - f_i|_1: ONE warp reduction per point, not one per neighbour
    for (int k = 0; k < K; ++k) {
      int j = __shfl_sync(SPF_FULL, my, k);
      float wk = __shfl_sync(SPF_FULL, w, k);
      if (j < 0) continue;
      float diff = feat[(size_t)j * 32 + lane] - fi;
      part += wk * fabsf(diff);
      if (grad) {
        float gsc = grad_scale * wk / norm / (float)N * sgnf(diff);
        if (gsc != 0.0f) atomicAdd(&grad[(size_t)j * 32 + lane], gsc);
        gi -= gsc;
      }
    }
    tv_i = warp_sum(part) / norm;
    if (grad && gi != 0.0f) atomicAdd(&grad[(size_t)i * 32 + lane], gi);
  }
  __shared__ float s_v[8];
  if (lane == 0) s_v[threadIdx.x >> 5] = active ? tv_i : 0.0f;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_v[w];
    atomicAdd(value, grad_scale * t / (float)N);
  }
}

extern "C" int spf_tv_fwd_bwd_range(const float* pts, const float* feat_g, const int32_t* self_pidx, int32_t N, int32_t K,
                                    int32_t first, int32_t count, float* value, float* grad, float scale, void* stream_) {
  if (!pts || !feat_g || !self_pidx || !value || first < 0 || count < 0) return SPF_ERR_INVALID;
  if (K > 32) return SPF_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream_;
  SPF_CUDA(cudaMemsetAsync(value, 0, sizeof(float), st), "tv memset");
  if (N <= 0 || count == 0) return SPF_OK;
  k_tv<<<(count + 7) / 8, 256, 0, st>>>(pts, feat_g, self_pidx, N, K, first, count, value, grad, scale);
  SPF_CHECK_LAUNCH("k_tv");
  return SPF_OK;
}

extern "C" int spf_tv_fwd_bwd(const float* pts, const float* feat_g, const int32_t* self_pidx, int32_t N, int32_t K,
                              float* value, float* grad, float grad_scale, void* stream_) {
  return spf_tv_fwd_bwd_range(pts, feat_g, self_pidx, N, K, 0, N, value, grad, grad_scale, stream_);
}

// ------------------------------------------------------------------------------------------------
// a14: VolSDFLoss (spurfies/model/loss.py:51-100) forward + its gradients w.r.t. the rendered colour and the
// compositing weights, fused (the torch version is ~70 tiny launches per step):
//   rgb_loss  = mean |rgb - gt|                                   (L1Loss, loss.py:29-32)
//   eikonal   = mean over valid samples of (|grad_theta|_2 - 1)^2 (loss.py:34-40; no gradient: SURVEY D8)
//   mask_loss = BCE(clip(sum_s w, 1e-3, 1 - 1e-3), mask)          (loss.py:80-84)
//   loss      = w_rgb rgb + w_eik eik + w_tv tv + w_local local + w_pseudo pseudo + mask
// Pass 1: one warp per ray, one thread per sample, per-block partial sums (fixed order -> run-to-run deterministic);
// pass 2: one block folds the partials in double and writes the terms.
// ------------------------------------------------------------------------------------------------
#define LOSS_THREADS 256
#define LOSS_MAX_BLOCKS 1024

__global__ void __launch_bounds__(LOSS_THREADS)
k_loss_partial(const float* __restrict__ rgb, const float* __restrict__ rgb_gt, const float* __restrict__ weights,
               const float* __restrict__ mask_gt, int mask_stride, const float* __restrict__ grad_theta,
               const uint8_t* __restrict__ valid, long long n, int R, int S, float w_rgb, float* __restrict__ d_rgb,
               float* __restrict__ d_weights, float* __restrict__ partial) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};   // sum |rgb - gt|, sum bce, sum (|g| - 1)^2, valid count
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float inv3R = 1.0f / (3.0f * (float)R), invR = 1.0f / (float)R;
  // rays: one warp per ray (coalesced reads of its S weights, coalesced writes of their gradient); lane c < 3 owns channel c
  const int lane_ = threadIdx.x & 31;
  const long long nwarps = stride >> 5;
  for (long long r = tid0 >> 5; r < R; r += nwarps) {
    if (lane_ < 3) {
      const float d = rgb[3 * r + lane_] - rgb_gt[3 * r + lane_];
      acc[0] += fabsf(d);
      if (d_rgb) d_rgb[3 * r + lane_] = w_rgb * sgnf(d) * inv3R;
    }
    if (weights) {
      float ws = 0.0f;
      for (int s = lane_; s < S; s += 32) ws += weights[r * S + s];
      ws = warp_sum(ws);
      const float y = mask_gt[r * mask_stride];
      const float x = fminf(fmaxf(ws, 1.0e-3f), 1.0f - 1.0e-3f);
      if (lane_ == 0) acc[1] += -(y * fmaxf(logf(x), -100.0f) + (1.0f - y) * fmaxf(logf(1.0f - x), -100.0f));
      if (d_weights) {
        // binary_cross_entropy backward (x - y) / max((1 - x) x, 1e-12), through the clip (pass-through inside [min, max])
        const bool inside = ws >= 1.0e-3f && ws <= 1.0f - 1.0e-3f;
        const float gx = inside ? (x - y) / fmaxf((1.0f - x) * x, 1.0e-12f) * invR : 0.0f;
        for (int s = lane_; s < S; s += 32) d_weights[r * S + s] = gx;
      }
    }
  }
  if (grad_theta) {
    for (long long i = tid0; i < n; i += stride) {
      if (valid[i]) {
        const float gx = grad_theta[3 * i], gy = grad_theta[3 * i + 1], gz = grad_theta[3 * i + 2];
        const float e = sqrtf(gx * gx + gy * gy + gz * gz) - 1.0f;
        acc[2] += e * e;
        acc[3] += 1.0f;
      }
    }
  }
  __shared__ float s_p[LOSS_THREADS / 32][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float v = warp_sum(acc[k]);
    if (lane == 0) s_p[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float v = 0.0f;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) v += s_p[w][threadIdx.x];
    partial[4 * blockIdx.x + threadIdx.x] = v;
  }
}

__global__ void k_loss_final(const float* __restrict__ partial, int nb, int R, int has_mask, const float* __restrict__ tv,
                             const float* __restrict__ local, const float* __restrict__ pseudo, float w_rgb, float w_eik,
                             float w_tv, float w_local, float w_pseudo, float* __restrict__ terms) {
  __shared__ double s_d[4][32];
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;   // 128 threads: warp k folds quantity k
  double v = 0.0;
  for (int b = lane; b < nb; b += 32) v += (double)partial[4 * b + k];
  s_d[k][lane] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double q[4];
    for (int j = 0; j < 4; ++j) {
      double a = 0.0;
      for (int l = 0; l < 32; ++l) a += s_d[j][l];
      q[j] = a;
    }
    const float rgb_loss = (float)(q[0] / (3.0 * (double)R));
    const float mask_loss = has_mask ? (float)(q[1] / (double)R) : 0.0f;
    const float eik = (float)(q[2] / (q[3] > 1.0 ? q[3] : 1.0));
    const float tvv = (tv && w_tv > 0.0f) ? tv[0] : 0.0f;
    const float loc = local ? local[0] : 0.0f;
    const float pse = (pseudo && w_pseudo > 0.0f) ? pseudo[0] : 0.0f;
    terms[0] = w_rgb * rgb_loss + w_eik * eik + w_tv * tvv + w_local * loc + w_pseudo * pse + mask_loss;
    terms[1] = rgb_loss; terms[2] = eik; terms[3] = tvv; terms[4] = mask_loss; terms[5] = loc; terms[6] = pse;
    terms[7] = (float)q[3];
  }
}

extern "C" size_t spf_loss_workspace_bytes(void) { return (size_t)LOSS_MAX_BLOCKS * 4 * sizeof(float); }

extern "C" int spf_volsdf_loss(const float* rgb, const float* rgb_gt, const float* weights, const float* mask_gt,
                               int32_t mask_stride, const float* grad_theta, const uint8_t* valid, int64_t n, int32_t R,
                               int32_t S, const float* tv, const float* local, const float* pseudo, float w_rgb, float w_eik,
                               float w_tv, float w_local, float w_pseudo, float* terms, float* d_rgb, float* d_weights,
                               void* workspace, size_t workspace_bytes, void* stream_) {
  if (!rgb || !rgb_gt || !terms || !workspace || R < 1) return SPF_ERR_INVALID;
  if (weights && (!mask_gt || S < 1)) return SPF_ERR_INVALID;
  if (grad_theta && !valid) return SPF_ERR_INVALID;
  if (workspace_bytes < spf_loss_workspace_bytes()) return SPF_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream_;
  const long long work = grad_theta && n > R ? n : R;
  long long nb = (work + LOSS_THREADS - 1) / LOSS_THREADS;
  if (nb > LOSS_MAX_BLOCKS) nb = LOSS_MAX_BLOCKS;
  k_loss_partial<<<(int)nb, LOSS_THREADS, 0, st>>>(rgb, rgb_gt, weights, mask_gt, mask_stride, grad_theta, valid, n, R, S,
                                                  w_rgb, d_rgb, d_weights, (float*)workspace);
  SPF_CHECK_LAUNCH("k_loss_partial");
  k_loss_final<<<1, 128, 0, st>>>((const float*)workspace, (int)nb, R, weights != nullptr, tv, local, pseudo, w_rgb, w_eik,
                                  w_tv, w_local, w_pseudo, terms);
  SPF_CHECK_LAUNCH("k_loss_final");
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
// a14: pseudo-point loss (pointneus_disent.py:765-780): L1 of the SDF at each hit ray's expected-depth point, mean over
// the rays that hit AND whose point has a neighbour; 1000 when rays hit but no point has one (the reference's 1000-filled
// rows), 0 when nothing hit.  One block; also writes the per-ray gradient factors the backward scales by the upstream
// gradient: u_sdf[r] = sign(sdf_r) ok_r / cnt (d loss / d sdf_r) and u_dist[r] = u_sdf[r] * (grad_r . dir_r) (d loss / d dist_r
// through x_r = cam + dist_r dir_r).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_pseudo_loss(const float* __restrict__ sdf, const float* __restrict__ grad,
                                                      const int* __restrict__ pidx, int K, const int* __restrict__ ray_nvalid,
                                                      const float* __restrict__ ray_dirs, int R, float* __restrict__ value,
                                                      float* __restrict__ u_sdf, float* __restrict__ u_dist) {
  __shared__ float s_sum[32];
  __shared__ int s_cnt[32], s_hit[32];
  __shared__ float s_inv;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float sum = 0.0f;
  int cnt = 0, hit = 0;
  for (int r = tid; r < R; r += blockDim.x) {
    const bool h = ray_nvalid[r] > 0;
    const bool ok = h && pidx[(size_t)r * K] >= 0;
    hit |= h ? 1 : 0;
    if (ok) { sum += fabsf(sdf[r]); ++cnt; }
  }
  sum = warp_sum(sum);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { cnt += __shfl_xor_sync(SPF_FULL, cnt, o); hit |= __shfl_xor_sync(SPF_FULL, hit, o); }
  if (lane == 0) { s_sum[warp] = sum; s_cnt[warp] = cnt; s_hit[warp] = hit; }
  __syncthreads();
  if (tid == 0) {
    float t = 0.0f;
    int c = 0, h = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t += s_sum[w]; c += s_cnt[w]; h |= s_hit[w]; }
    value[0] = c > 0 ? t / (float)c : (h ? 1000.0f : 0.0f);
    s_inv = c > 0 ? 1.0f / (float)c : 0.0f;
  }
  __syncthreads();
  const float inv = s_inv;
  for (int r = tid; r < R; r += blockDim.x) {
    const bool ok = ray_nvalid[r] > 0 && pidx[(size_t)r * K] >= 0;
    const float u = ok ? sgnf(sdf[r]) * inv : 0.0f;
    u_sdf[r] = u;
    if (u_dist)
      u_dist[r] = ok ? u * (grad[3 * r] * ray_dirs[3 * r] + grad[3 * r + 1] * ray_dirs[3 * r + 1] + grad[3 * r + 2] * ray_dirs[3 * r + 2]) : 0.0f;
  }
}

extern "C" int spf_pseudo_loss(const float* sdf, const float* grad, const int32_t* pidx, int32_t K, const int32_t* ray_nvalid,
                               const float* ray_dirs, int32_t R, float* value, float* u_sdf, float* u_dist, void* stream_) {
  if (!sdf || !pidx || !ray_nvalid || !value || !u_sdf || K < 1 || R < 1) return SPF_ERR_INVALID;
  if (u_dist && (!grad || !ray_dirs)) return SPF_ERR_INVALID;
  k_pseudo_loss<<<1, 1024, 0, (cudaStream_t)stream_>>>(sdf, grad, pidx, K, ray_nvalid, ray_dirs, R, value, u_sdf, u_dist);
  SPF_CHECK_LAUNCH("k_pseudo_loss");
  return SPF_OK;
}

// x[r] = cam + dist[r] * dir[r] (the expected-depth point of each ray, pointneus_disent.py:766-768)
__global__ void k_ray_points(const float* __restrict__ cam, const float* __restrict__ dirs, const float* __restrict__ dist,
                             int R, float* __restrict__ x) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float d = dist[r];
#pragma unroll
  for (int a = 0; a < 3; ++a) x[3 * r + a] = cam[a] + dirs[3 * r + a] * d;
}

extern "C" int spf_ray_points(const float* cam_loc, const float* ray_dirs, const float* dist, int32_t R, float* x,
                              void* stream_) {
  if (!cam_loc || !ray_dirs || !dist || !x) return SPF_ERR_INVALID;
  if (R <= 0) return SPF_OK;
  k_ray_points<<<(R + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(cam_loc, ray_dirs, dist, R, x);
  SPF_CHECK_LAUNCH("k_ray_points");
  return SPF_OK;
}
