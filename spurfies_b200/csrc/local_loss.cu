// local_loss.cu -- SURVEY 8(f4): the DTU feature-consistency ("local") loss that consumes the hot path's per-ray SDF.
//
//   find_surface_points   spurfies/model/pointneus_disent.py:586-612   first back-facing zero crossing per ray
//   get_local_loss        spurfies/feat_utils.py:377-451 (uncerts = None), called from pointneus_disent.py:727-763
//   idx_world2cam / idx_cam2img / normalize_for_grid_sample / get_in_range   feat_utils.py:43-77
//
// One warp per ray, one lane per feature channel (C = 32, the Vis-MVSNet width, feat_utils.py:355-357).  The loss of a
// ray depends on the network only through the scalar depth d of the crossing (the surface point is o + d * dir), so
// the derivative is carried FORWARD as one tangent (d/dd) through projection, bilinear sampling and the cosine
// correlation; the kernel emits the loss numerator and d numerator / d sdf at the two slots either side of the
// crossing.  Nothing is atomically accumulated: the host sums the per-ray numerators (fixed order).
#include "common.cuh"

#define LL_C 32

struct LLView {
  float gx, gy;      // normalised grid coordinate (after the clamp)
  float gxp, gyp;    // tangents d/dd (0 where the clamp is active)
  bool in_range;
};

// feat_utils.py:43-55 and 58-68 with the tangent of every step.  E row-major 4x4 (world -> camera), Kc row-major 4x4
// whose [:3,:3] are the intrinsics.
__device__ __forceinline__ LLView ll_project(const float* __restrict__ E, const float* __restrict__ Kc, const float pw[3],
                                             const float pwp[3], int H, int W) {
  float X[4], Xp[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    X[i] = E[i * 4 + 0] * pw[0] + E[i * 4 + 1] * pw[1] + E[i * 4 + 2] * pw[2] + E[i * 4 + 3];
    Xp[i] = E[i * 4 + 0] * pwp[0] + E[i * 4 + 1] * pwp[1] + E[i * 4 + 2] * pwp[2];
  }
  // idx_cam_homo / (idx_cam_homo[..., -1:, :] + 1e-9)            (:46)
  float w1 = X[3] + 1e-9f, w1p = Xp[3];
  float Y[4], Yp[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    Y[i] = X[i] / w1;
    Yp[i] = (Xp[i] * w1 - X[i] * w1p) / (w1 * w1);
  }
  // idx_cam_homo[..., :3, :] / (idx_cam_homo[..., 3:4, :] + 1e-9) (:52)
  float w2 = Y[3] + 1e-9f, w2p = Yp[3];
  float c[3], cp[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    c[i] = Y[i] / w2;
    cp[i] = (Yp[i] * w2 - Y[i] * w2p) / (w2 * w2);
  }
  float I[3], Ip[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    I[i] = Kc[i * 4 + 0] * c[0] + Kc[i * 4 + 1] * c[1] + Kc[i * 4 + 2] * c[2];
    Ip[i] = Kc[i * 4 + 0] * cp[0] + Kc[i * 4 + 1] * cp[1] + Kc[i * 4 + 2] * cp[2];
  }
  // idx_img_homo / (idx_img_homo[..., -1:, :] + 1e-9)             (:54)
  float z3 = I[2] + 1e-9f, z3p = Ip[2];
  float u = I[0] / z3, v = I[1] / z3;
  float up = (Ip[0] * z3 - I[0] * z3p) / (z3 * z3), vp = (Ip[1] * z3 - I[1] * z3p) / (z3 * z3);
  // grid / 2 (feat_utils.py:415), / [w, h] * 2 - 1, clamp(-1.1, 1.1) (:66-67)
  float gx = (u / 2.0f) / (float)W * 2.0f - 1.0f, gy = (v / 2.0f) / (float)H * 2.0f - 1.0f;
  LLView o;
  o.gxp = (gx >= -1.1f && gx <= 1.1f) ? up / (float)W : 0.0f;
  o.gyp = (gy >= -1.1f && gy <= 1.1f) ? vp / (float)H : 0.0f;
  o.gx = fminf(fmaxf(gx, -1.1f), 1.1f);
  o.gy = fminf(fmaxf(gy, -1.1f), 1.1f);
  if (gx != gx) o.gx = gx;  // clamp propagates NaN in torch
  if (gy != gy) o.gy = gy;
  o.in_range = (o.gx <= 1.0f) && (o.gx >= -1.0f) && (o.gy <= 1.0f) && (o.gy >= -1.0f);  // get_in_range (:71-77)
  return o;
}

// F.grid_sample(mode="bilinear", padding_mode="zeros", align_corners=False) of this lane's channel, with the tangent.
__device__ __forceinline__ void ll_sample(const float* __restrict__ feat, int64_t cs, int64_t ps, int H, int W,
                                          const LLView& g, int lane, float& f, float& fp) {
  float ix = ((g.gx + 1.0f) * (float)W - 1.0f) * 0.5f, iy = ((g.gy + 1.0f) * (float)H - 1.0f) * 0.5f;
  float ixp = g.gxp * (float)W * 0.5f, iyp = g.gyp * (float)H * 0.5f;
  f = 0.0f; fp = 0.0f;
  if (!(ix == ix) || !(iy == iy)) { f = ix + iy; return; }  // NaN in -> NaN out (as torch)
  float x0f = floorf(ix), y0f = floorf(iy);
  float wx = ix - x0f, wy = iy - y0f;
  int x0 = (int)x0f, y0 = (int)y0f;
  const float* base = feat + (int64_t)lane * cs;
  bool xa = (x0 >= 0 && x0 < W), xb = (x0 + 1 >= 0 && x0 + 1 < W), ya = (y0 >= 0 && y0 < H), yb = (y0 + 1 >= 0 && y0 + 1 < H);
  float f00 = (xa && ya) ? __ldg(base + ((int64_t)y0 * W + x0) * ps) : 0.0f;
  float f01 = (xb && ya) ? __ldg(base + ((int64_t)y0 * W + x0 + 1) * ps) : 0.0f;
  float f10 = (xa && yb) ? __ldg(base + ((int64_t)(y0 + 1) * W + x0) * ps) : 0.0f;
  float f11 = (xb && yb) ? __ldg(base + ((int64_t)(y0 + 1) * W + x0 + 1) * ps) : 0.0f;
  f = f00 * (1.0f - wx) * (1.0f - wy) + f01 * wx * (1.0f - wy) + f10 * (1.0f - wx) * wy + f11 * wx * wy;
  float fx = (f01 - f00) * (1.0f - wy) + (f11 - f10) * wy;
  float fy = (f10 - f00) * (1.0f - wx) + (f11 - f01) * wx;
  fp = fx * ixp + fy * iyp;
}

__global__ void __launch_bounds__(256)
k_local_loss(const float* __restrict__ sdf, const float* __restrict__ t, const float* __restrict__ cam_loc,
             const float* __restrict__ ray_dirs, int R, int Smax, const float* __restrict__ feat_ref,
             const float* __restrict__ feat_src, int64_t src_stride, int64_t cs, int64_t ps,
             const float* __restrict__ cam_ref, const float* __restrict__ cam_src, int m, int H, int W,
             const float* __restrict__ size, const float* __restrict__ center, float* __restrict__ num,
             int32_t* __restrict__ cross, float* __restrict__ d_surface, float* __restrict__ g0,
             float* __restrict__ g1) {
  int ray = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (ray >= R) return;
  const float* s = sdf + (int64_t)ray * Smax;
  const float* tt = t + (int64_t)ray * Smax;
  // first slot i with sdf[i] * sdf[i+1] < 0 and sdf[i+1] < sdf[i]; 1000 = no neighbour -> NaN (:587-595)
  int first = -1;
  for (int base = 0; base < Smax - 1 && first < 0; base += 32) {
    int i = base + lane;
    bool c = false;
    if (i + 1 < Smax) {
      float a = s[i], b = s[i + 1];
      if (a != 1000.0f && b != 1000.0f) c = (b * a < 0.0f) && (b < a);
    }
    unsigned bal = __ballot_sync(SPF_FULL, c);
    if (bal) first = base + __ffs(bal) - 1;
  }
  if (first < 0) {
    if (lane == 0) { num[ray] = 0.0f; cross[ray] = -1; d_surface[ray] = 0.0f; g0[ray] = 0.0f; g1[ray] = 0.0f; }
    return;
  }
  float s0 = s[first], s1 = s[first + 1], d0 = tt[first], d1 = tt[first + 1];
  float den = s0 - s1;
  float d = (s0 * d1 - s1 * d0) / den;  // :610
  float numer = 0.0f, dnum = 0.0f;
  if (m > 0) {
    float sz = size[0];
    float pw[3], pwp[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float dir = ray_dirs[(int64_t)ray * 3 + c];
      float p = cam_loc[c] + dir * d;            // pointneus_disent.py:745-748
      pw[c] = p / 2.0f * sz + center[c];         // feat_utils.py:402-404
      pwp[c] = dir / 2.0f * sz;
    }
    LLView v0 = ll_project(cam_ref, cam_ref + 16, pw, pwp, H, W);
    float a, ap;
    ll_sample(feat_ref, cs, ps, H, W, v0, lane, a, ap);
    float na2 = warp_sum(a * a), aap = warp_sum(a * ap);
    float na = sqrtf(na2);
    float nap = na > 0.0f ? aap / na : 0.0f;
    float nac = fmaxf(na, 1e-9f), nacp = (na >= 1e-9f) ? nap : 0.0f;   // .clamp(min=1e-9) (:432)
    for (int v = 0; v < m; ++v) {
      const float* cam = cam_src + (int64_t)v * 32;
      LLView vs = ll_project(cam, cam + 16, pw, pwp, H, W);
      float b, bp;
      ll_sample(feat_src + (int64_t)v * src_stride, cs, ps, H, W, vs, lane, b, bp);
      float ab = warp_sum(a * b), nb2 = warp_sum(b * b), apb = warp_sum(ap * b), abp = warp_sum(a * bp),
            bbp = warp_sum(b * bp);
      float nb = sqrtf(nb2);
      float nbp = nb > 0.0f ? bbp / nb : 0.0f;
      float nbc = fmaxf(nb, 1e-9f), nbcp = (nb >= 1e-9f) ? nbp : 0.0f;
      float corr = ab / nac / nbc;                                                     // :430-434
      float corrp = (apb + abp) / (nac * nbc) - corr * (nacp / nac + nbcp / nbc);
      float e = 1.0f - corr;
      float cl = fabsf(e);                                                             // :435
      float clp = (e > 0.0f ? -1.0f : (e < 0.0f ? 1.0f : 0.0f)) * corrp;
      bool keep = v0.in_range && vs.in_range && (cl < 0.5f);                           // :417-419, 437-438
      if (keep) { numer += cl; dnum += clp; }
    }
  }
  if (lane == 0) {
    num[ray] = numer;
    cross[ray] = first;
    d_surface[ray] = d;
    float q = dnum / (den * den);
    g0[ray] = q * s1 * (d0 - d1);   // d d / d s0
    g1[ray] = q * s0 * (d1 - d0);   // d d / d s1
  }
}

// d_sdf[r, s] = scale * (s == cross[r] ? g0[r] : s == cross[r] + 1 ? g1[r] : 0): every element is written
__global__ void k_local_loss_bwd(const int32_t* __restrict__ cross, const float* __restrict__ g0,
                                 const float* __restrict__ g1, const float* __restrict__ scale, int R, int Smax,
                                 float* __restrict__ d_sdf) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)R * Smax) return;
  int r = (int)(i / Smax), sl = (int)(i - (int64_t)r * Smax);
  int c = cross[r];
  float v = 0.0f;
  if (c >= 0) {
    if (sl == c) v = g0[r] * scale[0];
    else if (sl == c + 1) v = g1[r] * scale[0];
  }
  d_sdf[i] = v;
}

extern "C" int spf_local_loss_fwd(const float* sdf, const float* t, const float* cam_loc, const float* ray_dirs, int32_t R,
                                  int32_t Smax, const float* feat_ref, const float* feat_src, int64_t src_stride,
                                  int64_t chan_stride, int64_t pix_stride, int32_t channels, const float* cam_ref,
                                  const float* cam_src, int32_t m, int32_t H, int32_t W, const float* size,
                                  const float* center, float* num, int32_t* cross, float* d_surface, float* g0,
                                  float* g1, void* stream_) {
  if (!sdf || !t || !cam_loc || !ray_dirs || !num || !cross || !d_surface || !g0 || !g1 || Smax < 2) return SPF_ERR_INVALID;
  if (m < 0) return SPF_ERR_INVALID;
  if (m > 0) {
    if (!feat_ref || !feat_src || !cam_ref || !cam_src || !size || !center || H <= 0 || W <= 0) return SPF_ERR_INVALID;
    if (channels != LL_C) return SPF_ERR_UNSUPPORTED;
  }
  if (R <= 0) return SPF_OK;
  k_local_loss<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream_>>>(sdf, t, cam_loc, ray_dirs, R, Smax, feat_ref, feat_src,
                                                               src_stride, chan_stride, pix_stride, cam_ref, cam_src, m,
                                                               H, W, size, center, num, cross, d_surface, g0, g1);
  SPF_CHECK_LAUNCH("k_local_loss");
  return SPF_OK;
}

extern "C" int spf_local_loss_bwd(const int32_t* cross, const float* g0, const float* g1, const float* scale, int32_t R,
                                  int32_t Smax, float* d_sdf, void* stream_) {
  if (!cross || !g0 || !g1 || !scale || !d_sdf) return SPF_ERR_INVALID;
  if (R <= 0) return SPF_OK;
  int64_t n = (int64_t)R * Smax;
  k_local_loss_bwd<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(cross, g0, g1, scale, R, Smax, d_sdf);
  SPF_CHECK_LAUNCH("k_local_loss_bwd");
  return SPF_OK;
}
