// grid.cu -- voxel grid build, occupancy mask / per-ray slots, warp-cooperative kNN, compaction.
//
// B200-first redesign of torch_knnquery/src/knnquery.cu (six 1-thread-per-item kernels over dense
// X*Y*Z int grids + a [max_o,P] table):
//   * points are counting-sorted by reference-geometry voxel into ONE float4 array (x,y,z,id), so a
//     query streams whole cells with coalesced 16-byte loads; the three z-neighbour cells of an
//     (x,y) column are contiguous, so the 27-cell neighbourhood is 9 contiguous segments;
//   * one WARP per query: 32 candidates per step, ballot of the survivors, and a warp-resident sorted
//     top-K (one key per lane, 64-bit key = d2 bits << 32 | point id) -- results are the K smallest by
//     (d2, id), emitted sorted, independent of insertion order (the reference's are order dependent,
//     knnquery.cu:280-299);
//   * no P / max_o caps and no curand reservoir (knnquery.cu:69-79, 158-165): nothing is dropped;
//   * everything is launched on the caller's stream and checked (the reference uses the legacy
//     default stream and never checks, knnquery.cu:339).
// Compiled with -fmad=false: fp32 arithmetic matches the reference op for op; the one contraction
// nvcc applies in the reference (d2, knnquery.cu:281) is written out with explicit fmaf.
#include <stdio.h>
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// grid build: count -> scan -> fill -> dilate
// ------------------------------------------------------------------------------------------------
__global__ void k_grid_count(GridDev g, const float* __restrict__ pts, int n, int* __restrict__ cell_of,
                             int* __restrict__ counts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cx, cy, cz;
  int c = voxel_of(g, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], cx, cy, cz);  // knnquery.cu:45-50
  cell_of[i] = c;
  if (c >= 0) atomicAdd(&counts[c], 1);
}

// single-block exclusive scan of counts[G] -> cell_start[G+1]; also stats (occupied, max count, total)
__global__ void k_grid_scan(const int* __restrict__ counts, int G, int* __restrict__ cell_start,
                            int* __restrict__ cursor, int* __restrict__ stats) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  int occ = 0, mx = 0;
  for (int base = 0; base < G; base += blockDim.x) {
    int i = base + tid;
    int v = i < G ? counts[i] : 0;
    occ += v > 0;
    mx = max(mx, v);
    int inc = warp_scan_incl_i(v, lane);
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int w = lane < (blockDim.x >> 5) ? s_warp[lane] : 0;
      int winc = warp_scan_incl_i(w, lane);
      s_warp[lane] = winc - w;  // exclusive offset of each warp
    }
    __syncthreads();
    int excl = s_carry + s_warp[wid] + inc - v;
    if (i < G) {
      cell_start[i] = excl;
      cursor[i] = excl;
    }
    __syncthreads();
    if (tid == blockDim.x - 1) s_carry = excl + v;
    __syncthreads();
  }
  // reduce stats
  for (int o = 16; o > 0; o >>= 1) {
    occ += __shfl_xor_sync(SPF_FULL, occ, o);
    mx = max(mx, __shfl_xor_sync(SPF_FULL, mx, o));
  }
  __shared__ int s_occ[32], s_mx[32];
  if (lane == 0) { s_occ[wid] = occ; s_mx[wid] = mx; }
  __syncthreads();
  if (tid == 0) {
    int o = 0, m = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { o += s_occ[w]; m = max(m, s_mx[w]); }
    cell_start[G] = s_carry;
    stats[0] = o; stats[1] = m; stats[2] = s_carry; stats[3] = 0;
  }
}

__global__ void k_grid_fill(const float* __restrict__ pts, int n, const int* __restrict__ cell_of,
                            int* __restrict__ cursor, float4* __restrict__ sorted) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = cell_of[i];
  if (c < 0) return;
  int pos = atomicAdd(&cursor[c], 1);
  sorted[pos] = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], __int_as_float(i));
}

// hit[v] = any occupied o with o - k/2 <= v < o + (k+1)/2 per axis (knnquery.cu:110-116), gather form
__global__ void k_grid_dilate(GridDev g, int G, uint8_t* __restrict__ hit) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= G) return;
  int z = v % g.dz, y = (v / g.dz) % g.dy, x = v / (g.dz * g.dy);
  int h = 0;
  for (int ox = max(0, x - (g.kx + 1) / 2 + 1); ox <= min(g.dx - 1, x + g.kx / 2) && !h; ++ox)
    for (int oy = max(0, y - (g.ky + 1) / 2 + 1); oy <= min(g.dy - 1, y + g.ky / 2) && !h; ++oy)
      for (int oz = max(0, z - (g.kz + 1) / 2 + 1); oz <= min(g.dz - 1, z + g.kz / 2); ++oz) {
        int o = ox * (g.dy * g.dz) + oy * g.dz + oz;
        if (g.cell_start[o + 1] > g.cell_start[o]) { h = 1; break; }
      }
  hit[v] = (uint8_t)h;
}

// search-grid cell of a point: only points inside the reference grid are inserted (the reference ignores the others)
__global__ void k_search_count(GridDev g, const float* __restrict__ pts, int n, int* __restrict__ cell_of,
                               int* __restrict__ counts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
  int cx, cy, cz;
  int c = voxel_of(g, x, y, z, cx, cy, cz);
  if (c >= 0) {
    int fx = (int)floorf(__fdiv_rn(__fsub_rn(x, g.sx), g.fcell));
    int fy = (int)floorf(__fdiv_rn(__fsub_rn(y, g.sy), g.fcell));
    int fz = (int)floorf(__fdiv_rn(__fsub_rn(z, g.sz), g.fcell));
    c = (fx < 0 || fx >= g.fdx || fy < 0 || fy >= g.fdy || fz < 0 || fz >= g.fdz) ? -1 : fx * (g.fdy * g.fdz) + fy * g.fdz + fz;
  }
  cell_of[i] = c;
  if (c >= 0) atomicAdd(&counts[c], 1);
}

extern "C" size_t spf_grid_workspace_bytes(int32_t n_points, int32_t n_cells) {
  return sizeof(int) * ((size_t)n_points + 2 * (size_t)n_cells + 64);
}

extern "C" int spf_grid_build(const spf_grid* g, const float* points, int32_t* cell_start, float* sorted,
                              uint8_t* hit, int32_t* stats, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!g || !points || !cell_start || !sorted || !hit || !stats || !workspace) return SPF_ERR_INVALID;
  int n = g->n_points, G = g->n_cells;
  if (G != g->dim[0] * g->dim[1] * g->dim[2] || G <= 0 || n < 0) return SPF_ERR_INVALID;
  if (workspace_bytes < spf_grid_workspace_bytes(n, G)) return SPF_ERR_WORKSPACE;
  int* cell_of = (int*)workspace;
  int* counts = cell_of + n;
  int* cursor = counts + G;
  GridDev d = to_dev(g);
  d.cell_start = cell_start;
  d.sorted = reinterpret_cast<const float4*>(sorted);
  d.hit = hit;
  SPF_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)G, st), "grid_build memset");
  if (n > 0) {
    k_grid_count<<<(n + 255) / 256, 256, 0, st>>>(d, points, n, cell_of, counts);
    SPF_CHECK_LAUNCH("k_grid_count");
  }
  k_grid_scan<<<1, 1024, 0, st>>>(counts, G, cell_start, cursor, stats);
  SPF_CHECK_LAUNCH("k_grid_scan");
  if (n > 0) {
    k_grid_fill<<<(n + 255) / 256, 256, 0, st>>>(points, n, cell_of, cursor, reinterpret_cast<float4*>(sorted));
    SPF_CHECK_LAUNCH("k_grid_fill");
  }
  k_grid_dilate<<<(G + 255) / 256, 256, 0, st>>>(d, G, hit);
  SPF_CHECK_LAUNCH("k_grid_dilate");
  return SPF_OK;
}

extern "C" int spf_grid_build_search(const spf_grid* g, const float* points, int32_t* search_cell_start, float* search_sorted,
                                     void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!g || !points || !search_cell_start || !search_sorted || !workspace) return SPF_ERR_INVALID;
  const int n = g->n_points;
  const long long G = (long long)g->search_dim[0] * g->search_dim[1] * g->search_dim[2];
  if (G <= 0 || G >= (1ll << 31) || !(g->search_cell > 0.0f)) return SPF_ERR_INVALID;
  if (workspace_bytes < spf_grid_workspace_bytes(n, (int)G)) return SPF_ERR_WORKSPACE;
  int* cell_of = (int*)workspace;
  int* counts = cell_of + n;
  int* cursor = counts + G;
  int* stats = cursor + G;   // 64 spare ints at the end of the workspace
  GridDev d = to_dev(g);
  SPF_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)G, st), "grid_build_search memset");
  if (n > 0) {
    k_search_count<<<(n + 255) / 256, 256, 0, st>>>(d, points, n, cell_of, counts);
    SPF_CHECK_LAUNCH("k_search_count");
  }
  k_grid_scan<<<1, 1024, 0, st>>>(counts, (int)G, search_cell_start, cursor, stats);
  SPF_CHECK_LAUNCH("k_grid_scan");
  if (n > 0) {
    k_grid_fill<<<(n + 255) / 256, 256, 0, st>>>(points, n, cell_of, cursor, reinterpret_cast<float4*>(search_sorted));
    SPF_CHECK_LAUNCH("k_grid_fill");
  }
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
// warp-cooperative kNN of one query against the cells within Chebyshev distance L of its voxel
// (knnquery.cu:263-303).  Returns this lane's key (lanes 0..K-1 hold the sorted result).
// ------------------------------------------------------------------------------------------------
#define KEY_NONE 0xffffffffffffffffull

__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src) {
  unsigned lo = __shfl_sync(SPF_FULL, (unsigned)v, src);
  unsigned hi = __shfl_sync(SPF_FULL, (unsigned)(v >> 32), src);
  return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long shfl_up_u64(unsigned long long v) {
  unsigned lo = __shfl_up_sync(SPF_FULL, (unsigned)v, 1);
  unsigned hi = __shfl_up_sync(SPF_FULL, (unsigned)(v >> 32), 1);
  return ((unsigned long long)hi << 32) | lo;
}

__device__ __forceinline__ unsigned long long knn_warp(const GridDev& g, float qx, float qy, float qz, int K,
                                                       float r2, int lane) {
  // candidate cells: the 27 reference voxels around the query (knnquery.cu:263-277), or -- same result, ~3.4x fewer
  // candidates -- the 27 search cells (edge >= radius) that cover the radius ball
  const bool fine = g.use_search != 0;
  const float vx = fine ? g.fcell : g.vx, vy = fine ? g.fcell : g.vy, vz = fine ? g.fcell : g.vz;
  const int dx = fine ? g.fdx : g.dx, dy = fine ? g.fdy : g.dy, dz = fine ? g.fdz : g.dz;
  const int* __restrict__ cstart = fine ? g.fcell_start : g.cell_start;
  const float4* __restrict__ cand = fine ? g.fsorted : g.sorted;
  int fx = (int)floorf(__fdiv_rn(__fsub_rn(qx, g.sx), vx));
  int fy = (int)floorf(__fdiv_rn(__fsub_rn(qy, g.sy), vy));
  int fz = (int)floorf(__fdiv_rn(__fsub_rn(qz, g.sz), vz));
  const int L = fine ? 1 : (g.kx + 1) / 2 - 1;  // the reference uses kernel_size[0] on all axes (knnquery.cu:263)
  unsigned long long key = KEY_NONE, thr = KEY_NONE;
  const int z0 = max(0, fz - L), z1 = min(dz - 1, fz + L);
  if (z0 > z1) return key;
  // (Skipping the cells of the 27 that the radius ball cannot reach was measured and rejected: with cell edge ~ radius
  // only ~20 % of the corner columns can be skipped, and the per-column distance test costs more instructions than the
  // shorter scan saves: k_knn_slots 0.27 -> 0.32 ms.  Likewise rejected: fetching the nine columns' bounds in one round
  // trip and streaming the chained segments 32 candidates per step with the next step prefetched -- the kernel is
  // issue-bound on the insertion shuffles (SM pipe 79 % busy), not latency-bound: 0.273 -> 0.307 ms.  And: visiting the
  // query's own (x, y) column first so that the threshold tightens early -- fewer insertions, but the index arithmetic and
  // the divergent skip of out-of-range columns cost more: 0.285 -> 0.326 ms.)
  for (int cx = max(0, fx - L); cx <= min(dx - 1, fx + L); ++cx)
    for (int cy = max(0, fy - L); cy <= min(dy - 1, fy + L); ++cy) {
      const int base = cx * (dy * dz) + cy * dz;
      const int beg = cstart[base + z0], end = cstart[base + z1 + 1];
      for (int j0 = beg; j0 < end; j0 += 32) {
        int j = j0 + lane;
        bool pass = false;
        unsigned long long ck = KEY_NONE;
        if (j < end) {
          float4 c = cand[j];
          float xv = __fsub_rn(c.x, qx), yv = __fsub_rn(c.y, qy), zv = __fsub_rn(c.z, qz);
          // knnquery.cu:281 as nvcc contracts it: FMUL y*y ; FFMA x,x ; FFMA z,z
          float d2 = __fmaf_rn(zv, zv, __fmaf_rn(xv, xv, __fmul_rn(yv, yv)));
          pass = (r2 == 0.0f) || (d2 <= r2);  // knnquery.cu:282
          ck = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(c.w);
          pass = pass && (ck < thr);
        }
        unsigned m = __ballot_sync(SPF_FULL, pass);
        while (m) {
          int src = __ffs(m) - 1;
          m &= m - 1;
          unsigned long long k = shfl_u64(ck, src);
          if (k < thr) {  // warp-uniform
            unsigned long long up = shfl_up_u64(key);
            if (key > k) key = (lane == 0 || up < k) ? k : up;
            thr = shfl_u64(key, K - 1);
          }
        }
      }
    }
  return key;
}

// the search grid may replace the reference voxels only when its 27 cells cover the radius ball AND the ball lies inside
// the 27 reference voxels (radius <= every voxel edge): then both candidate sets contain exactly the same in-radius points
static inline GridDev to_dev_query(const spf_grid* g, float radius2) {
  GridDev d = to_dev(g);
  if (g->search_sorted && g->search_cell_start && radius2 > 0.0f) {
    const float r = sqrtf(radius2);
    if (r * 1.0005f <= g->search_cell && r <= g->vsize[0] && r <= g->vsize[1] && r <= g->vsize[2]) d.use_search = 1;
  }
  return d;
}

// ------------------------------------------------------------------------------------------------
// THREAD-per-query kNN (K <= 8, radius > 0: every call of the hot path).  The warp-per-query scan above spends its issue
// slots on the serial top-K insertion (six shuffles per surviving candidate, ~25 insertions per query on a DTU-shaped
// cloud where ~66 of the ~250 scanned candidates lie inside the radius) and on 9 mostly half-empty 32-lane steps.  Here a
// lane owns a query:
//   * the lanes of a warp are consecutive slots of a ray (or consecutive query points), so they walk the same few cell
//     columns and their candidate loads hit the same L1 lines;
//   * a candidate costs the distance arithmetic, ONE compare against a running limit (min(radius^2, d2 of the current
//     K-th best)) and an unconditional 8-byte store of its key into the lane's shared-memory list, whose fill count
//     advances only if the candidate passed: no divergent branch per candidate;
//   * the sorted top-K lives in registers; lists are merged into it ("drain") only when some lane's list is nearly full
//     and once at the end -- warp-synchronously, so the branchy insertion network runs ~2-4 times x list length per
//     warp instead of once per survivor per query, and every drain tightens the limit;
//   * a block first compacts its valid queries (slots below the ray's count / points inside the dilated occupancy), so
//     warps are dense.
// The result is the K smallest by (d2, id), as before: the selection is a total order, so it cannot depend on the scan
// order -- bit-identical to knn_warp (tests/test_gpu_knn.py compares both with the oracle and with each other).
// ------------------------------------------------------------------------------------------------
#define KT_THREADS 128   // measured on the 4096-ray step: 64-thread blocks 0.235 ms, 128: 0.209 ms; 2 / 4 slots per thread
#define KT_CAP 24        // (256 / 512 slots per block, dense rounds as in k_knn_points_t): 0.232 / 0.292 ms
#define KT_UNROLL 4
#define KT_KR 8

__device__ __forceinline__ void kt_insert(unsigned long long (&key)[KT_KR], unsigned long long k) {
#pragma unroll
  for (int j = KT_KR - 1; j > 0; --j) {
    const bool g1 = key[j - 1] > k, g0 = key[j] > k;
    key[j] = g1 ? key[j - 1] : (g0 ? k : key[j]);
  }
  if (key[0] > k) key[0] = k;
}

// all 32 lanes: merge each lane's list into its sorted keys, empty the lists, tighten the limits
__device__ __forceinline__ void kt_drain(unsigned long long (&key)[KT_KR], const unsigned long long* list, int& n, float& lim) {
  const int nmax = __reduce_max_sync(SPF_FULL, n);
  for (int i = 0; i < nmax; ++i) {
    const unsigned long long k = i < n ? list[i * KT_THREADS] : KEY_NONE;
    if (k < key[KT_KR - 1]) kt_insert(key, k);
  }
  n = 0;
  // a candidate farther than the current 8th best cannot enter (equal distance can: the point id decides, in kt_insert)
  if (key[KT_KR - 1] != KEY_NONE) lim = __uint_as_float((unsigned)(key[KT_KR - 1] >> 32));
}

// Called by all 32 lanes of a warp (inactive lanes scan nothing).  list = this lane's column of the block's list array.
__device__ __forceinline__ void knn_thread(const GridDev& g, bool active, float qx, float qy, float qz, float r2,
                                           unsigned long long* list, unsigned long long (&key)[KT_KR]) {
  const bool fine = g.use_search != 0;
  const float vx = fine ? g.fcell : g.vx, vy = fine ? g.fcell : g.vy, vz = fine ? g.fcell : g.vz;
  const int dx = fine ? g.fdx : g.dx, dy = fine ? g.fdy : g.dy, dz = fine ? g.fdz : g.dz;
  const int* __restrict__ cstart = fine ? g.fcell_start : g.cell_start;
  const float4* __restrict__ cand = fine ? g.fsorted : g.sorted;
  const int fx = (int)floorf(__fdiv_rn(__fsub_rn(qx, g.sx), vx));
  const int fy = (int)floorf(__fdiv_rn(__fsub_rn(qy, g.sy), vy));
  const int fz = (int)floorf(__fdiv_rn(__fsub_rn(qz, g.sz), vz));
  const int L = fine ? 1 : (g.kx + 1) / 2 - 1;  // the reference uses kernel_size[0] on all axes (knnquery.cu:263)
#pragma unroll
  for (int j = 0; j < KT_KR; ++j) key[j] = KEY_NONE;
  const int z0 = max(0, fz - L), z1 = min(dz - 1, fz + L);
  if (z0 > z1) active = false;
  int n = 0;
  float lim = r2;
  const float INF = __int_as_float(0x7f800000);
  for (int ox = -L; ox <= L; ++ox)
    for (int oy = -L; oy <= L; ++oy) {
      const int cx = fx + ox, cy = fy + oy;
      int beg = 1, len = 0;   // an empty segment reads element 0 (index beg + (-1)) and discards it
      if (active && cx >= 0 && cx < dx && cy >= 0 && cy < dy) {
        const int base = cx * (dy * dz) + cy * dz;
        const int b = cstart[base + z0], e = cstart[base + z1 + 1];
        if (e > b) { beg = b; len = e - b; }
      }
      const int maxlen = __reduce_max_sync(SPF_FULL, len);
      const int last = len - 1;
      for (int o = 0; o < maxlen; o += KT_UNROLL) {
        float4 c[KT_UNROLL];
#pragma unroll
        for (int u = 0; u < KT_UNROLL; ++u) c[u] = cand[beg + min(o + u, last)];   // past the segment: re-reads, discarded
#pragma unroll
        for (int u = 0; u < KT_UNROLL; ++u) {
          const float xv = __fsub_rn(c[u].x, qx), yv = __fsub_rn(c[u].y, qy), zv = __fsub_rn(c[u].z, qz);
          // knnquery.cu:281 as nvcc contracts it: FMUL y*y ; FFMA x,x ; FFMA z,z
          const float d2 = __fmaf_rn(zv, zv, __fmaf_rn(xv, xv, __fmul_rn(yv, yv)));
          *reinterpret_cast<uint2*>(list + n * KT_THREADS) = make_uint2((unsigned)__float_as_int(c[u].w), __float_as_uint(d2));
          // knnquery.cu:282 (radius) and the running K-th-best bound in one compare
          n += (d2 <= lim && o + u <= last) ? 1 : 0;
        }
        if (__any_sync(SPF_FULL, n > KT_CAP - KT_UNROLL)) kt_drain(key, list, n, lim);
      }
    }
  kt_drain(key, list, n, lim);
}

// block-level compaction of the valid work items: returns the flat id this thread processes (-1: none).  All threads call.
__device__ __forceinline__ int kt_compact(bool valid, int flat) {
  __shared__ int s_item[KT_THREADS];
  __shared__ int s_wsum[KT_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned m = __ballot_sync(SPF_FULL, valid);
  if (lane == 0) s_wsum[wid] = __popc(m);
  __syncthreads();
  int off = 0, total = 0;
#pragma unroll
  for (int w = 0; w < KT_THREADS / 32; ++w) {
    const int c = s_wsum[w];
    if (w < wid) off += c;
    total += c;
  }
  if (valid) s_item[off + __popc(m & ((1u << lane) - 1))] = flat;
  __syncthreads();
  return tid < total ? s_item[tid] : -1;
}

__global__ void __launch_bounds__(KT_THREADS) k_knn_slots_t(GridDev g, const float* __restrict__ sample_loc,
                                                            const int* __restrict__ n_slots, int R, int Smax, int K, float r2,
                                                            int* __restrict__ pidx, int* __restrict__ ray_nvalid) {
  __shared__ unsigned long long s_list[KT_CAP * KT_THREADS];
  const long long total = (long long)R * Smax;
  const long long flat0 = (long long)blockIdx.x * KT_THREADS;
  const long long mine = flat0 + threadIdx.x;
  bool valid = false;
  if (mine < total) {
    const int r = (int)(mine / Smax), s = (int)(mine - (long long)r * Smax);
    valid = s < n_slots[r];
    if (!valid)
      for (int k = 0; k < K; ++k) pidx[mine * K + k] = -1;
  }
  const int item = kt_compact(valid, threadIdx.x);
  if (__all_sync(SPF_FULL, item < 0)) return;
  const long long w = flat0 + (item < 0 ? 0 : item);
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (item >= 0) { qx = sample_loc[3 * w]; qy = sample_loc[3 * w + 1]; qz = sample_loc[3 * w + 2]; }
  unsigned long long key[KT_KR];
  knn_thread(g, item >= 0, qx, qy, qz, r2, s_list + threadIdx.x, key);
  if (item >= 0) {
#pragma unroll
    for (int k = 0; k < KT_KR; ++k)
      if (k < K) pidx[w * K + k] = key[k] == KEY_NONE ? -1 : (int)(unsigned)(key[k] & 0xffffffffull);
    if (key[0] != KEY_NONE) atomicAdd(&ray_nvalid[(int)(w / Smax)], 1);
  }
}

// Point queries are sparse (coarse ray samples: a quarter inside the dilated occupancy; SDF grids: a few per cent), so a
// block compacts KT_PITEMS x 128 consecutive points and works through the survivors in dense rounds of 128 -- with one
// point per thread most blocks kept a single half-empty warp alive and the SM ran 9 warps (0.27 ms for the step's coarse
// pass against 0.15 ms of the warp-per-query kernel).
#define KT_PITEMS 8
__global__ void __launch_bounds__(KT_THREADS) k_knn_points_t(GridDev g, const float* __restrict__ q, long long Q, int K, float r2,
                                                             int* __restrict__ pidx, const int* __restrict__ skip) {
  __shared__ unsigned long long s_list[KT_CAP * KT_THREADS];
  __shared__ int s_item[KT_PITEMS * KT_THREADS];
  __shared__ int s_wsum[KT_PITEMS * (KT_THREADS / 32)];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long flat0 = (long long)blockIdx.x * (KT_PITEMS * KT_THREADS);
  const bool skipped = skip && *skip;
  // validity of this thread's KT_PITEMS points (point it * 128 + tid of the block), -1 rows for the others
  unsigned vbits = 0, ball[KT_PITEMS];
#pragma unroll
  for (int it = 0; it < KT_PITEMS; ++it) {
    const long long mine = flat0 + it * KT_THREADS + tid;
    bool valid = false;
    if (mine < Q) {
      int cx, cy, cz;
      const int v = skipped ? -1 : voxel_of(g, q[3 * mine], q[3 * mine + 1], q[3 * mine + 2], cx, cy, cz);
      valid = v >= 0 && g.hit[v];
      if (!valid)
        for (int k = 0; k < K; ++k) pidx[mine * K + k] = -1;
    }
    ball[it] = __ballot_sync(SPF_FULL, valid);
    vbits |= (unsigned)valid << it;
    if (lane == 0) s_wsum[it * (KT_THREADS / 32) + wid] = __popc(ball[it]);
  }
  __syncthreads();
  int total = 0;
#pragma unroll
  for (int it = 0; it < KT_PITEMS; ++it) {
    int off = total;
#pragma unroll
    for (int w = 0; w < KT_THREADS / 32; ++w) {
      const int c = s_wsum[it * (KT_THREADS / 32) + w];
      if (w < wid) off += c;
      total += c;
    }
    if ((vbits >> it) & 1) s_item[off + __popc(ball[it] & ((1u << lane) - 1))] = it * KT_THREADS + tid;
  }
  __syncthreads();
  for (int base = 0; base < total; base += KT_THREADS) {
    if (base + (wid << 5) >= total) break;              // warp-uniform: this warp has no work in this round or later
    const int item = base + tid < total ? s_item[base + tid] : -1;
    const long long w = flat0 + (item < 0 ? 0 : item);
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (item >= 0) { qx = q[3 * w]; qy = q[3 * w + 1]; qz = q[3 * w + 2]; }
    unsigned long long key[KT_KR];
    knn_thread(g, item >= 0, qx, qy, qz, r2, s_list + tid, key);
    if (item >= 0) {
#pragma unroll
      for (int k = 0; k < KT_KR; ++k)
        if (k < K) pidx[w * K + k] = key[k] == KEY_NONE ? -1 : (int)(unsigned)(key[k] & 0xffffffffull);
    }
  }
}

// 0 = automatic, 1 = always the warp-per-query kernels, 2 = the thread-per-query kernels wherever they apply (K <= 8,
// radius > 0).  Automatic = thread-per-query for RAY SLOTS (consecutive slots of a ray walk the same cell columns, so a
// warp's candidate loads share L1 lines: k_knn_slots 0.274 -> 0.205 ms on the 4096-ray DTU-shaped step, 13.1 -> 6.9 ms
// per full eval image) and for very large POINT batches (SDF grids); warp-per-query for the other point queries and for
// clouds with very dense voxels (spf_grid.dense_cloud: garden-shaped 1 M points, ~1300 candidates per query, 32 lanes in
// 32 different cells make every candidate load 32 L1 wavefronts: 0.77 -> 1.20 ms with a thread per query).
static int g_knn_algo = 0;
extern "C" int spf_knn_set_algo(int32_t algo) {
  if (algo < 0 || algo > 2) return SPF_ERR_INVALID;
  g_knn_algo = algo;
  return SPF_OK;
}
static inline bool knn_use_thread_kernels(const spf_grid* g, int K, float radius2, bool ray_slots, long long Q = 0) {
  if (K > KT_KR || !(radius2 > 0.0f) || g_knn_algo == 1) return false;
  if (g_knn_algo == 2) return true;
  if (g->dense_cloud) return false;
  // point queries: only the large batches (SDF grids for marching cubes: the ~2 M in-occupancy points of a 16 M-point
  // chunk, consecutive along z: 512^3 volume 35.0 -> 14.6 ms); the step's coarse pass (0.5 M points, 0.15 ms) has too
  // few valid points to fill the machine with one thread per query, and a 2 M-point eval chunk is a draw
  return ray_slots || Q >= (1ll << 20);
}

// ------------------------------------------------------------------------------------------------
// mask + slots, one warp per ray (knnquery.cu:171-221, knnquery.py:208-231)
// ------------------------------------------------------------------------------------------------
__global__ void k_mask_slots(GridDev g, const float* __restrict__ raypos, int R, int D, int Smax,
                             int* __restrict__ slot_sample, float* __restrict__ sample_loc,
                             int* __restrict__ n_slots) {
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= R) return;
  int cum = 0;
  for (int base = 0; base < D; base += 32) {
    int d = base + lane;
    bool h = false;
    float x = 0, y = 0, z = 0;
    if (d < D) {
      const float* p = raypos + ((size_t)r * D + d) * 3;
      x = p[0]; y = p[1]; z = p[2];
      int cx, cy, cz;
      int v = voxel_of(g, x, y, z, cx, cy, cz);
      h = v >= 0 && g.hit[v];
    }
    unsigned m = __ballot_sync(SPF_FULL, h);
    int slot = cum + __popc(m & ((1u << lane) - 1));
    if (h && slot < Smax) {
      size_t s = (size_t)r * Smax + slot;
      slot_sample[s] = d;
      sample_loc[3 * s] = x; sample_loc[3 * s + 1] = y; sample_loc[3 * s + 2] = z;
    }
    cum += __popc(m);
  }
  int ns = min(cum, Smax);
  for (int s = ns + lane; s < Smax; s += 32) {
    size_t o = (size_t)r * Smax + s;
    slot_sample[o] = -1;
    sample_loc[3 * o] = 0.f; sample_loc[3 * o + 1] = 0.f; sample_loc[3 * o + 2] = 0.f;
  }
  if (lane == 0) n_slots[r] = ns;
}

// one warp per (ray, slot)
__global__ void k_knn_slots(GridDev g, const float* __restrict__ sample_loc, const int* __restrict__ n_slots,
                            int R, int Smax, int K, float r2, int* __restrict__ pidx, int* __restrict__ ray_nvalid) {
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= (long long)R * Smax) return;
  int r = (int)(w / Smax), s = (int)(w - (long long)r * Smax);
  unsigned long long key = KEY_NONE;
  if (s < n_slots[r]) {
    const float* q = sample_loc + 3 * w;
    key = knn_warp(g, q[0], q[1], q[2], K, r2, lane);
    if (lane == 0 && key != KEY_NONE) atomicAdd(&ray_nvalid[r], 1);
  }
  if (lane < K) pidx[w * K + lane] = key == KEY_NONE ? -1 : (int)(unsigned)(key & 0xffffffffull);
}

// point queries: mask + kNN fused, each warp owns `ppw` consecutive points (ppw = 32 for huge, mostly masked-out
// batches such as SDF grids; smaller when there are too few points to fill the machine with 32 per warp)
__global__ void k_knn_points(GridDev g, const float* __restrict__ q, long long Q, int K, float r2,
                             int* __restrict__ pidx, int ppw, const int* __restrict__ skip) {
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  long long i = w * ppw + lane;
  if (w * ppw >= Q) return;
  const bool skipped = skip && *skip;   // device-side predicate (spf_knn_points_pred): every point is "outside"
  bool h = false;
  float x = 0, y = 0, z = 0;
  if (lane < ppw && i < Q && skipped) {
    for (int k = 0; k < K; ++k) pidx[i * K + k] = -1;
  } else if (lane < ppw && i < Q) {
    x = q[3 * i]; y = q[3 * i + 1]; z = q[3 * i + 2];
    int cx, cy, cz;
    int v = voxel_of(g, x, y, z, cx, cy, cz);
    h = v >= 0 && g.hit[v];
    if (!h)
      for (int k = 0; k < K; ++k) pidx[i * K + k] = -1;
  }
  unsigned m = __ballot_sync(SPF_FULL, h);
  while (m) {
    int src = __ffs(m) - 1;
    m &= m - 1;
    float qx = __shfl_sync(SPF_FULL, x, src), qy = __shfl_sync(SPF_FULL, y, src), qz = __shfl_sync(SPF_FULL, z, src);
    unsigned long long key = knn_warp(g, qx, qy, qz, K, r2, lane);
    if (lane < K) pidx[(w * ppw + src) * K + lane] = key == KEY_NONE ? -1 : (int)(unsigned)(key & 0xffffffffull);
  }
}

__global__ void k_mask_points(GridDev g, const float* __restrict__ q, long long Q, int* __restrict__ mask) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Q) return;
  int cx, cy, cz;
  int v = voxel_of(g, q[3 * i], q[3 * i + 1], q[3 * i + 2], cx, cy, cz);
  mask[i] = v >= 0 ? (int)g.hit[v] : 0;  // knnquery.cu:192-195
}

extern "C" int spf_mask_slots(const spf_grid* g, const float* raypos, int32_t R, int32_t D, int32_t Smax,
                              int32_t* slot_sample, float* sample_loc, int32_t* n_slots, void* stream_) {
  if (!g || !raypos || !slot_sample || !sample_loc || !n_slots || D <= 0 || Smax <= 0) return SPF_ERR_INVALID;
  if (R <= 0) return SPF_OK;
  const int wpb = 8;
  k_mask_slots<<<(R + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream_>>>(to_dev(g), raypos, R, D, Smax,
                                                                             slot_sample, sample_loc, n_slots);
  SPF_CHECK_LAUNCH("k_mask_slots");
  return SPF_OK;
}

extern "C" int spf_knn_slots(const spf_grid* g, const float* sample_loc, const int32_t* n_slots, int32_t R,
                             int32_t Smax, int32_t K, float radius2, int32_t* pidx, int32_t* ray_nvalid,
                             void* stream_) {
  if (!g || !sample_loc || !n_slots || !pidx || !ray_nvalid) return SPF_ERR_INVALID;
  if (K < 1 || K > 20) return SPF_ERR_INVALID;  // knnquery.py:184
  if (R <= 0) return SPF_OK;
  cudaStream_t st = (cudaStream_t)stream_;
  SPF_CUDA(cudaMemsetAsync(ray_nvalid, 0, sizeof(int) * (size_t)R, st), "knn_slots memset");
  if (knn_use_thread_kernels(g, K, radius2, true)) {
    const long long total = (long long)R * Smax;
    k_knn_slots_t<<<(unsigned)((total + KT_THREADS - 1) / KT_THREADS), KT_THREADS, 0, st>>>(
        to_dev_query(g, radius2), sample_loc, n_slots, R, Smax, K, radius2, pidx, ray_nvalid);
    SPF_CHECK_LAUNCH("k_knn_slots_t");
    return SPF_OK;
  }
  const int wpb = 8;
  long long warps = (long long)R * Smax;
  k_knn_slots<<<(unsigned)((warps + wpb - 1) / wpb), wpb * 32, 0, st>>>(to_dev_query(g, radius2), sample_loc, n_slots, R, Smax, K,
                                                                        radius2, pidx, ray_nvalid);
  SPF_CHECK_LAUNCH("k_knn_slots");
  return SPF_OK;
}

extern "C" int spf_knn_points_pred(const spf_grid* g, const float* q, int64_t Q, int32_t K, float radius2, int32_t* pidx,
                                   const int32_t* skip, void* stream_) {
  if (K < 1 || K > 20) return SPF_ERR_INVALID;
  if (Q <= 0) return SPF_OK;
  if (!g || !q || !pidx) return SPF_ERR_INVALID;
  if (knn_use_thread_kernels(g, K, radius2, false, Q)) {
    const long long per_block = (long long)KT_PITEMS * KT_THREADS;
    k_knn_points_t<<<(unsigned)((Q + per_block - 1) / per_block), KT_THREADS, 0, (cudaStream_t)stream_>>>(
        to_dev_query(g, radius2), q, Q, K, radius2, pidx, skip);
    SPF_CHECK_LAUNCH("k_knn_points_t");
    return SPF_OK;
  }
  const int wpb = 8;
  // enough warps to fill the machine several times over (148 SMs x 64 resident warps), at most 32 points per warp
  // (measured on the step's coarse pass, 0.5 M points: x2 -> 0.184 ms, x8 -> 0.150 ms, x32 / x128 -> 0.195 ms)
  int ppw = 32;
  const long long want_warps = (long long)spf_num_sms() * 64 * 8;
  while (ppw > 1 && (Q + ppw - 1) / ppw < want_warps) ppw >>= 1;
  long long warps = (Q + ppw - 1) / ppw;
  k_knn_points<<<(unsigned)((warps + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream_>>>(to_dev_query(g, radius2), q, Q, K,
                                                                                            radius2, pidx, ppw, skip);
  SPF_CHECK_LAUNCH("k_knn_points");
  return SPF_OK;
}

extern "C" int spf_knn_points(const spf_grid* g, const float* q, int64_t Q, int32_t K, float radius2, int32_t* pidx,
                              void* stream_) {
  return spf_knn_points_pred(g, q, Q, K, radius2, pidx, nullptr, stream_);
}

extern "C" int spf_mask_points(const spf_grid* g, const float* q, int64_t Q, int32_t* mask, void* stream_) {
  if (Q <= 0) return SPF_OK;
  if (!g || !q || !mask) return SPF_ERR_INVALID;
  k_mask_points<<<(unsigned)((Q + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(to_dev(g), q, Q, mask);
  SPF_CHECK_LAUNCH("k_mask_points");
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
// compaction of valid slots: 3-phase scan (block counts -> scan of block counts -> scatter)
// ------------------------------------------------------------------------------------------------
#define CMP_THREADS 256
#define CMP_ITEMS 4
#define CMP_TILE (CMP_THREADS * CMP_ITEMS)

__global__ void k_cmp_count(const int* __restrict__ pidx, long long n, int K, int* __restrict__ block_sums) {
  long long base = (long long)blockIdx.x * CMP_TILE;
  int c = 0;
#pragma unroll
  for (int it = 0; it < CMP_ITEMS; ++it) {
    long long i = base + it * CMP_THREADS + threadIdx.x;
    if (i < n) c += pidx[i * K] >= 0;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(SPF_FULL, c, o);
  __shared__ int s[CMP_THREADS / 32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < CMP_THREADS / 32; ++w) t += s[w];
    block_sums[blockIdx.x] = t;
  }
}

__global__ void k_cmp_scan_blocks(int* __restrict__ block_sums, int nb, int* __restrict__ count) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += blockDim.x) {
    int i = base + tid;
    int v = i < nb ? block_sums[i] : 0;
    int inc = warp_scan_incl_i(v, lane);
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int w = lane < (blockDim.x >> 5) ? s_warp[lane] : 0;
      int winc = warp_scan_incl_i(w, lane);
      s_warp[lane] = winc - w;
    }
    __syncthreads();
    int excl = s_carry + s_warp[wid] + inc - v;
    if (i < nb) block_sums[i] = excl;
    __syncthreads();
    if (tid == blockDim.x - 1) s_carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) *count = s_carry;
}

__global__ void k_cmp_scatter(const int* __restrict__ pidx, long long n, int K, const int* __restrict__ block_offs,
                              int* __restrict__ list) {
  // items are assigned so that output order is ascending in i: thread t owns items [t*ITEMS, t*ITEMS+ITEMS)
  long long base = (long long)blockIdx.x * CMP_TILE + (long long)threadIdx.x * CMP_ITEMS;
  int f[CMP_ITEMS];
  int c = 0;
#pragma unroll
  for (int it = 0; it < CMP_ITEMS; ++it) {
    long long i = base + it;
    f[it] = (i < n) && (pidx[i * K] >= 0);
    c += f[it];
  }
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = warp_scan_incl_i(c, lane);
  __shared__ int s[CMP_THREADS / 32];
  if (lane == 31) s[wid] = inc;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < wid; ++w) woff += s[w];
  int pos = block_offs[blockIdx.x] + woff + inc - c;
#pragma unroll
  for (int it = 0; it < CMP_ITEMS; ++it)
    if (f[it]) list[pos++] = (int)(base + it);
}

extern "C" size_t spf_compact_workspace_bytes(int64_t n) {
  return sizeof(int) * (size_t)((n + CMP_TILE - 1) / CMP_TILE + 64);
}

extern "C" int spf_compact_valid(const int32_t* pidx, int64_t n, int32_t K, int32_t* list, int32_t* count,
                                 void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!pidx || !list || !count || !workspace || K < 1) return SPF_ERR_INVALID;
  if (n >= (1ll << 31)) return SPF_ERR_UNSUPPORTED;
  if (workspace_bytes < spf_compact_workspace_bytes(n)) return SPF_ERR_WORKSPACE;
  if (n <= 0) {
    SPF_CUDA(cudaMemsetAsync(count, 0, sizeof(int), st), "compact memset");
    return SPF_OK;
  }
  int nb = (int)((n + CMP_TILE - 1) / CMP_TILE);
  int* bs = (int*)workspace;
  k_cmp_count<<<nb, CMP_THREADS, 0, st>>>(pidx, n, K, bs);
  SPF_CHECK_LAUNCH("k_cmp_count");
  k_cmp_scan_blocks<<<1, 1024, 0, st>>>(bs, nb, count);
  SPF_CHECK_LAUNCH("k_cmp_scan_blocks");
  k_cmp_scatter<<<nb, CMP_THREADS, 0, st>>>(pidx, n, K, bs, list);
  SPF_CHECK_LAUNCH("k_cmp_scatter");
  return SPF_OK;
}

// ------------------------------------------------------------------------------------------------
static char g_err[256] = "";
extern "C" void spf_set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
}
extern "C" const char* spf_last_cuda_error(void) { return g_err; }
extern "C" const char* spf_version(void) { return "spurfies_b200 0.1.0 (sm_100a)"; }
