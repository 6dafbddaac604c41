/*
 * oracle/knn_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C, single-threaded restatement of the reference voxel-grid kNN
 * (torch_knnquery @ 947957e).  Each function cites the reference lines it follows.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg may load this library.
 *
 * Semantics reproduced:
 *   - voxel of p: (int)floorf((p - shift) / vsize), fp32, true division
 *       (torch_knnquery/src/knnquery.cu:45-47, 185-187, 254-256)
 *   - occupied voxels + kernel_size dilation            (knnquery.cu:85-120)
 *   - per-sample mask = dilated occupancy of its voxel  (knnquery.cu:171-196)
 *   - per-ray slots: first Smax mask-hit samples        (knnquery.py:208-231, knnquery.cu:199-221)
 *   - per-slot kNN over the (kernel_size[0]+1)/2 - 1 Chebyshev shell of voxels,
 *     d2 = fma(z,z,fma(x,x,y*y)) exactly as nvcc contracts knnquery.cu:278-281
 *     (checked on sm_100a SASS: FMUL y*y; FFMA x,x; FFMA z,z), keep K smallest with
 *     d2 <= radius2 (radius2 == 0 disables the limit)   (knnquery.cu:263-303)
 *
 * Where the reference is order-dependent (replace-farthest with first-seen ties,
 * atomics arrival order, curand reservoir when a cap binds: knnquery.cu:69-79,
 * 155-165, 280-299) this oracle defines the canonical answer: the K smallest by
 * (d2, point id), emitted sorted by that key, no per-voxel / occupied-voxel caps.
 * spf_oracle_grid_stats reports the occupancy so callers can tell when the
 * reference's caps (P, max_o) would have made ITS answer random.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC knn_oracle.c -o _build/libknn_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  float shift[3];
  float vsize[3];
  int dim[3];
  int ks[3];
  int n_points;
  const float* pts; /* borrowed */
  int* cell_start;  /* [G+1] CSR over voxels */
  int* cell_pts;    /* [n_in_grid] point ids, ascending inside each voxel */
  uint8_t* hit;     /* [G] dilated occupancy */
  int G;
} grid_t;

static inline int voxel_of(const grid_t* g, const float* p, int c[3]) {
  /* knnquery.cu:45-49 */
  for (int a = 0; a < 3; ++a) {
    c[a] = (int)floorf((p[a] - g->shift[a]) / g->vsize[a]);
    if (c[a] < 0 || c[a] >= g->dim[a]) return -1;
  }
  return c[0] * (g->dim[1] * g->dim[2]) + c[1] * g->dim[2] + c[2]; /* knnquery.cu:50 */
}

void* spf_oracle_grid_create(const float* pts, int n, const float* shift,
                             const float* vsize, const int* dim, const int* ks) {
  grid_t* g = (grid_t*)calloc(1, sizeof(grid_t));
  memcpy(g->shift, shift, 12);
  memcpy(g->vsize, vsize, 12);
  memcpy(g->dim, dim, 12);
  memcpy(g->ks, ks, 12);
  g->n_points = n;
  g->pts = pts;
  g->G = dim[0] * dim[1] * dim[2];
  g->cell_start = (int*)calloc((size_t)g->G + 1, sizeof(int));
  g->hit = (uint8_t*)calloc((size_t)g->G, 1);
  int* cell_of = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int c[3];
  for (int i = 0; i < n; ++i) {
    cell_of[i] = voxel_of(g, pts + 3 * i, c);
    if (cell_of[i] >= 0) g->cell_start[cell_of[i] + 1]++;
  }
  for (int v = 0; v < g->G; ++v) g->cell_start[v + 1] += g->cell_start[v];
  g->cell_pts = (int*)malloc(sizeof(int) * (size_t)(g->cell_start[g->G] > 0 ? g->cell_start[g->G] : 1));
  int* cur = (int*)malloc(sizeof(int) * (size_t)(g->G > 0 ? g->G : 1));
  memcpy(cur, g->cell_start, sizeof(int) * (size_t)g->G);
  for (int i = 0; i < n; ++i)
    if (cell_of[i] >= 0) g->cell_pts[cur[cell_of[i]]++] = i;
  /* dilation: knnquery.cu:110-116, range [c - k/2, c + (k+1)/2) clipped */
  for (int x = 0; x < dim[0]; ++x)
    for (int y = 0; y < dim[1]; ++y)
      for (int z = 0; z < dim[2]; ++z) {
        int v = x * dim[1] * dim[2] + y * dim[2] + z;
        if (g->cell_start[v + 1] == g->cell_start[v]) continue;
        for (int xx = (x - ks[0] / 2 > 0 ? x - ks[0] / 2 : 0);
             xx < (x + (ks[0] + 1) / 2 < dim[0] ? x + (ks[0] + 1) / 2 : dim[0]); ++xx)
          for (int yy = (y - ks[1] / 2 > 0 ? y - ks[1] / 2 : 0);
               yy < (y + (ks[1] + 1) / 2 < dim[1] ? y + (ks[1] + 1) / 2 : dim[1]); ++yy)
            for (int zz = (z - ks[2] / 2 > 0 ? z - ks[2] / 2 : 0);
                 zz < (z + (ks[2] + 1) / 2 < dim[2] ? z + (ks[2] + 1) / 2 : dim[2]); ++zz)
              g->hit[xx * dim[1] * dim[2] + yy * dim[2] + zz] = 1;
      }
  free(cur);
  free(cell_of);
  return g;
}

void spf_oracle_grid_destroy(void* h) {
  grid_t* g = (grid_t*)h;
  if (!g) return;
  free(g->cell_start);
  free(g->cell_pts);
  free(g->hit);
  free(g);
}

/* out[0]=#occupied voxels, out[1]=max points in a voxel, out[2]=#points inside grid */
void spf_oracle_grid_stats(void* h, int* out) {
  grid_t* g = (grid_t*)h;
  int occ = 0, mx = 0;
  for (int v = 0; v < g->G; ++v) {
    int c = g->cell_start[v + 1] - g->cell_start[v];
    if (c > 0) occ++;
    if (c > mx) mx = c;
  }
  out[0] = occ;
  out[1] = mx;
  out[2] = g->cell_start[g->G];
}

/* knnquery.cu:171-196 : mask[i] = dilated occupancy of sample i's voxel (0 outside) */
void spf_oracle_mask(void* h, const float* q, int64_t n, int32_t* mask) {
  grid_t* g = (grid_t*)h;
  int c[3];
  for (int64_t i = 0; i < n; ++i) {
    int v = voxel_of(g, q + 3 * i, c);
    mask[i] = v >= 0 ? g->hit[v] : 0;
  }
}

typedef struct {
  float d2;
  int id;
} cand_t;

static inline int cand_less(float d2a, int ida, float d2b, int idb) {
  return d2a < d2b || (d2a == d2b && ida < idb);
}

/* knnquery.cu:224-308 for one query; writes K ids sorted by (d2,id), -1 padded.
 * Also returns the number of candidates scanned (for the roofline numerator). */
static int knn_one(const grid_t* g, const float* q, int K, float radius2, int32_t* out,
                   float* out_d2) {
  cand_t best[32];
  int nb = 0, scanned = 0;
  int f[3];
  for (int a = 0; a < 3; ++a) f[a] = (int)floorf((q[a] - g->shift[a]) / g->vsize[a]);
  int L = (g->ks[0] + 1) / 2 - 1; /* layers 0..L: knnquery.cu:263 */
  for (int x = (-f[0] > -L ? -f[0] : -L); x < (g->dim[0] - f[0] < L + 1 ? g->dim[0] - f[0] : L + 1); ++x)
    for (int y = (-f[1] > -L ? -f[1] : -L); y < (g->dim[1] - f[1] < L + 1 ? g->dim[1] - f[1] : L + 1); ++y)
      for (int z = (-f[2] > -L ? -f[2] : -L); z < (g->dim[2] - f[2] < L + 1 ? g->dim[2] - f[2] : L + 1); ++z) {
        int v = (f[0] + x) * g->dim[1] * g->dim[2] + (f[1] + y) * g->dim[2] + (f[2] + z);
        for (int j = g->cell_start[v]; j < g->cell_start[v + 1]; ++j) {
          int id = g->cell_pts[j];
          const float* p = g->pts + 3 * id;
          float xv = p[0] - q[0], yv = p[1] - q[1], zv = p[2] - q[2];
          float d2 = fmaf(zv, zv, fmaf(xv, xv, yv * yv)); /* knnquery.cu:281 as contracted by nvcc */
          scanned++;
          if (!(radius2 == 0.0f || d2 <= radius2)) continue; /* knnquery.cu:282 */
          if (nb == K && !cand_less(d2, id, best[K - 1].d2, best[K - 1].id)) continue;
          int pos = nb < K ? nb : K - 1;
          while (pos > 0 && cand_less(d2, id, best[pos - 1].d2, best[pos - 1].id)) {
            best[pos] = best[pos - 1];
            pos--;
          }
          best[pos].d2 = d2;
          best[pos].id = id;
          if (nb < K) nb++;
        }
      }
  for (int k = 0; k < K; ++k) {
    out[k] = k < nb ? best[k].id : -1;
    if (out_d2) out_d2[k] = k < nb ? best[k].d2 : -1.0f;
  }
  return scanned;
}

/*
 * Dense ray query.  raypos [R,D,3].  Outputs (all dense over the R input rays):
 *   slot_sample [R,Smax]  index d of the sample occupying the slot, -1 = empty
 *   sample_loc  [R,Smax,3] (0 for empty slots)            (knnquery.cu:199-221)
 *   pidx        [R,Smax,K] sorted by (d2,id), -1 padded   (knnquery.cu:224-308)
 *   ray_mask1   [R] any mask-hit sample                   (knnquery.py:212)
 *   ray_mask2   [R] any neighbour found                   (knnquery.py:272-280)
 * returns total candidates scanned.
 */
int64_t spf_oracle_query(void* h, const float* raypos, int R, int D, int Smax, int K,
                         float radius2, int32_t* slot_sample, float* sample_loc, int32_t* pidx,
                         float* pd2, int8_t* ray_mask1, int8_t* ray_mask2) {
  grid_t* g = (grid_t*)h;
  int64_t scanned = 0;
  int c[3];
  for (int r = 0; r < R; ++r) {
    int cum = 0, any = 0, anynb = 0;
    for (int s = 0; s < Smax; ++s) {
      slot_sample[(int64_t)r * Smax + s] = -1;
      for (int a = 0; a < 3; ++a) sample_loc[((int64_t)r * Smax + s) * 3 + a] = 0.0f;
      for (int k = 0; k < K; ++k) {
        pidx[((int64_t)r * Smax + s) * K + k] = -1;
        if (pd2) pd2[((int64_t)r * Smax + s) * K + k] = -1.0f;
      }
    }
    for (int d = 0; d < D; ++d) {
      const float* q = raypos + ((int64_t)r * D + d) * 3;
      int v = voxel_of(g, q, c);
      int m = v >= 0 ? g->hit[v] : 0;
      if (!m) continue;
      any = 1;
      cum++;
      if (cum > Smax) continue; /* knnquery.py:231 */
      int64_t s = (int64_t)r * Smax + (cum - 1);
      slot_sample[s] = d;
      memcpy(sample_loc + s * 3, q, 12);
      scanned += knn_one(g, q, K, radius2, pidx + s * K, pd2 ? pd2 + s * K : 0);
      if (pidx[s * K] >= 0) anynb = 1;
    }
    ray_mask1[r] = (int8_t)any;
    ray_mask2[r] = (int8_t)anynb;
  }
  return scanned;
}

/* Brute force: K smallest (d2,id) with d2 <= radius2 over ALL points (no grid).
 * Follows torch_knnquery/test/test_queries.py:22-44 (cdist + topk + radius mask) with the
 * kernel's d2 arithmetic; used to check that the grid restatement finds the true sets. */
void spf_oracle_brute(const float* pts, int n, const float* q, int64_t nq, int K, float radius2,
                      int32_t* out) {
  cand_t best[32];
  for (int64_t i = 0; i < nq; ++i) {
    int nb = 0;
    const float* qq = q + 3 * i;
    for (int id = 0; id < n; ++id) {
      const float* p = pts + 3 * id;
      float xv = p[0] - qq[0], yv = p[1] - qq[1], zv = p[2] - qq[2];
      float d2 = fmaf(zv, zv, fmaf(xv, xv, yv * yv));
      if (!(radius2 == 0.0f || d2 <= radius2)) continue;
      if (nb == K && !cand_less(d2, id, best[K - 1].d2, best[K - 1].id)) continue;
      int pos = nb < K ? nb : K - 1;
      while (pos > 0 && cand_less(d2, id, best[pos - 1].d2, best[pos - 1].id)) {
        best[pos] = best[pos - 1];
        pos--;
      }
      best[pos].d2 = d2;
      best[pos].id = id;
      if (nb < K) nb++;
    }
    for (int k = 0; k < K; ++k) out[i * K + k] = k < nb ? best[k].id : -1;
  }
}
