// tile_engine.cuh -- CTA-pair (cta_group::2), warp-specialised tile engine for the per-pair MLP chains (mlp_tc2.cu).
//
// Why: in the first-generation kernels (mlp_tc.cu) every CTA streams a whole 256x256 bf16 weight image (128 KB) from L2
// for every layer of every 128-row tile, and MMA / epilogue / weight load run back to back: ~3.5 us per tile-layer of
// which 1.04 us is tensor time, and 5.3 TB/s of L2->SM weight traffic (44 % of the ~12 TB/s L2 cap) at only 41 % of the
// tensor peak.  Here:
//   * two CTAs of a cluster (one TPC) form a pair; the leader issues tcgen05.mma.cta_group::2 with M = 256: each CTA
//     supplies its own 128 rows of A and HALF of the weight rows (N/2) -> each CTA loads 64 KB per layer instead of 128;
//   * each CTA keeps TWO row tiles (X, Y) in flight against the same resident weights (two 256-column fp32 accumulators
//     = all 512 TMEM columns): another 2x less weight traffic (32 KB per tile-layer), and the epilogue of one tile
//     overlaps the MMAs of the other;
//   * weights flow through a ring of 16 KB slots (one k-block of this CTA's half) filled by a producer thread with bulk
//     copies (TMA engine) on mbarriers and released by tcgen05.commit, so the next layer's weights arrive while the
//     current layer is still being multiplied;
//   * roles: warps 0-7 = epilogue of tile X, warps 8-15 = epilogue of tile Y (TMEM lane = tile row; within a group
//     warps 0-3 / 4-7 take the low / high 128 accumulator columns), so the two tiles' epilogues hide each other's
//     latencies; warp 16 lane 0 = weight producer, warp 17 lane 0 = MMA issuer (leader CTA) or weight-arrival relay
//     (peer CTA).
//
// Synchronisation (all mbarriers live at the same shared-memory offsets in both CTAs):
//   w_full[s]   (1 + tx)  TMA landed this CTA's half of the chunk in slot s
//   w_peer[s]   (1)       leader only: the peer's half landed (relay thread arrives remotely)
//   w_empty[s]  (1)       commit-multicast: all MMAs that read slot s have completed (both CTAs may refill)
//   a_ready[t]  (2)       leader only: both CTAs finished writing tile t's A operand and draining its accumulator
//   acc_full[t] (1)       commit-multicast: tile t's accumulator holds the layer's result
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace eng {
using namespace tc;

// ---- optional timeline instrumentation (development only: build with SPF_TIMELINE=1) --------------------
#ifdef SPF_TIMELINE
// Events are appended to a small ring in the CTA's spare shared memory (a shared-memory atomic + clock read: tens of
// cycles, so the instrumented kernel keeps its shape) and copied out by CTA 0 in teardown().
__device__ unsigned long long g_tl[4 * 8192];
__device__ unsigned int g_tl_n;
constexpr int TL_OFF = 220672;        // == SMEM_BYTES of the normal build (checked below)
constexpr int TL_MAX = 700;           // 16 B per event
__device__ __forceinline__ void tl(int ev, int t, int l) {
  extern __shared__ __align__(1024) uint8_t tl_smem[];
  if (blockIdx.x != 0 || (ev >= 10 && (threadIdx.x & 31) != 0)) return;
  unsigned* cnt = reinterpret_cast<unsigned*>(tl_smem + TL_OFF);
  const unsigned i = atomicAdd(cnt, 1u);
  if (i < TL_MAX) {
    unsigned long long* e = reinterpret_cast<unsigned long long*>(tl_smem + TL_OFF + 16 + 16 * i);
    e[0] = (unsigned long long)ev | ((unsigned long long)t << 8) | ((unsigned long long)l << 16);
    e[1] = clock64();
  }
}
__device__ __forceinline__ void tl_init() {
  extern __shared__ __align__(1024) uint8_t tl_smem[];
  if (threadIdx.x == 0) *reinterpret_cast<unsigned*>(tl_smem + TL_OFF) = 0u;
}
__device__ __forceinline__ void tl_dump() {
  extern __shared__ __align__(1024) uint8_t tl_smem[];
  if (blockIdx.x != 0) return;
  unsigned n = *reinterpret_cast<unsigned*>(tl_smem + TL_OFF);
  if (n > TL_MAX) n = TL_MAX;
  for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long* e = reinterpret_cast<const unsigned long long*>(tl_smem + TL_OFF + 16 + 16 * i);
    g_tl[4 * i] = e[0] & 0xff; g_tl[4 * i + 1] = (e[0] >> 8) & 0xff; g_tl[4 * i + 2] = (e[0] >> 16) & 0xff; g_tl[4 * i + 3] = e[1];
  }
  if (threadIdx.x == 0) g_tl_n = n;
}
#define TL(ev, t, l) tl(ev, t, l)
#define TL_INIT() tl_init()
#define TL_DUMP() tl_dump()
__device__ int g_dbg_mode;   // 0 normal, 1 epilogue skips TMEM loads, 2 epilogue loads but skips math + smem stores
#ifdef SPF_DBGMODE
#define DBG_MODE g_dbg_mode
#else
#define DBG_MODE 0
#endif
#else
#define TL(ev, t, l)
#define TL_INIT()
#define TL_DUMP()
#define DBG_MODE 0
#endif

constexpr int EPI_THREADS = 256;                     // per tile: 8 warps (TMEM lane quarter x column half)
constexpr int N_EPI_WARPS = 16;                      // warps 0-7: tile X, warps 8-15: tile Y
constexpr int WARP_PRODUCER = 16, WARP_MMA = 17;
constexpr int THREADS = 576;
constexpr int NSLOT = 5;
constexpr int SLOT_BYTES = 16384;
constexpr int A_BYTES = 65536;                       // 128 rows x 256 bf16, 4 k-blocks of 16 KB
constexpr int OFF_A = 0;                             // A_X, A_Y
constexpr int OFF_W = 2 * A_BYTES;
constexpr int OFF_PART = OFF_W + NSLOT * SLOT_BYTES; // 2 tiles x 256 floats
constexpr int OFF_BIAS = OFF_PART + 2048;            // 5 x 256 floats (per-kernel use)
constexpr int OFF_CHAIN = OFF_BIAS + 5120;           // the layer table (struct Chain)
constexpr int OFF_BAR = OFF_CHAIN + 256;
#ifdef SPF_TIMELINE
constexpr int SMEM_BYTES = OFF_BAR + 256 + 16 + 16 * 700;   // + the event ring
static_assert(TL_OFF == OFF_BAR + 256 && SMEM_BYTES <= 232448, "timeline ring placement");
#else
constexpr int SMEM_BYTES = OFF_BAR + 256;            // 220 672 <= 232 448
#endif
constexpr int MAX_LAYERS = 8;

struct Layer {
  const uint8_t* img;   // packed image of W [N][K] (packing.py): k-block stride N*128 bytes
  int nkb;              // k-blocks of 64
  int ksteps;           // K / 16 actually multiplied
  int N;                // output columns (multiple of 16, <= 256); each CTA holds N/2 weight rows
  int fmt;              // operand formats (umma.cuh): FMT_F16 = forward layers, FMT_BF16 = gradient layers
};
struct Chain {
  Layer L[MAX_LAYERS];
  int n;
};

struct Bars {
  uint64_t* w_full;   // [NSLOT]
  uint64_t* w_empty;  // [NSLOT]
  uint64_t* w_peer;   // [NSLOT]
  uint64_t* acc_full; // [2]
  uint64_t* a_ready;  // [2]
  uint32_t* tmem_ptr;
};
__device__ __forceinline__ Bars carve_bars(uint8_t* smem) {
  Bars b;
  b.w_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  b.w_empty = b.w_full + NSLOT;
  b.w_peer = b.w_empty + NSLOT;
  b.acc_full = b.w_peer + NSLOT;
  b.a_ready = b.acc_full + 2;
  b.tmem_ptr = reinterpret_cast<uint32_t*>(b.a_ready + 2);
  return b;
}

// common prologue: barrier init, TMEM allocation (512 columns, both CTAs), cluster-wide visibility. Returns TMEM base.
__device__ __forceinline__ uint32_t setup(uint8_t* smem, const Bars& b) {
  TL_INIT();
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(b.w_full + s, 1); mbar_init(b.w_empty + s, 1); mbar_init(b.w_peer + s, 1); }
    for (int t = 0; t < 2; ++t) { mbar_init(b.acc_full + t, 1); mbar_init(b.a_ready + t, 2); }
    fence_mbar_init();
  }
  if ((threadIdx.x >> 5) == 0) tmem_alloc2(b.tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  return *b.tmem_ptr;
}
__device__ __forceinline__ void teardown(uint32_t tmem) {
  tc_fence_before();
  __syncthreads();
  TL_DUMP();
  cluster_sync_all();   // the peer may still be signalling our barriers / the leader's MMAs still read our smem
  if ((threadIdx.x >> 5) == 0) tmem_dealloc2(tmem, 512);
}

// ---- weight producer (one thread per CTA) ----------------------------------------------------------
__device__ __forceinline__ void producer_loop(const Chain& ch, int n_iter, uint32_t rank, uint8_t* smem, const Bars& b) {
  uint32_t slot = 0, use = 0;
  for (int it = 0; it < n_iter; ++it)
    for (int l = 0; l < ch.n; ++l) {
      const Layer L = ch.L[l];
      const uint32_t bytes = (uint32_t)(L.N >> 1) * 128u;
      for (int kb = 0; kb < L.nkb; ++kb) {
        if (use > 0) mbar_wait(b.w_empty + slot, (use - 1) & 1);
        mbar_expect_tx(b.w_full + slot, bytes);
        bulk_g2s(smem + OFF_W + slot * SLOT_BYTES, L.img + (size_t)kb * ((size_t)L.N * 128) + (size_t)rank * bytes, bytes,
                 b.w_full + slot);
        if (++slot == NSLOT) { slot = 0; ++use; }
      }
    }
}
// ---- peer CTA: tell the leader when our half of each chunk has landed ----------------------------------
__device__ __forceinline__ void relay_loop(const Chain& ch, int n_iter, const Bars& b) {
  uint32_t slot = 0, use = 0;
  for (int it = 0; it < n_iter; ++it)
    for (int l = 0; l < ch.n; ++l)
      for (int kb = 0; kb < ch.L[l].nkb; ++kb) {
        mbar_wait(b.w_full + slot, use & 1);
        mbar_arrive_remote(b.w_peer + slot, 0);
        if (++slot == NSLOT) { slot = 0; ++use; }
      }
}
// ---- leader CTA: MMA issue.  Run by ALL 32 lanes of the MMA warp (warp-uniform control flow, every lane polls the
// barriers); one elected lane issues.  Descriptors are a precomputed base plus a small immediate: the issue path
// must stay far below the 128 cycles one M256 N256 K16 instruction takes to execute.
__device__ __forceinline__ void mma_loop(const Chain& ch, int n_iter, uint8_t* smem, const Bars& b, uint32_t tmem) {
  const uint64_t adesc0 = smem_desc_sw128(smem_u32(smem + OFF_A));
  const uint64_t bdesc0 = smem_desc_sw128(smem_u32(smem + OFF_W));
  uint32_t slot0 = 0, use0 = 0;
  uint32_t ar_par = 0;   // bit t = parity of a_ready[t]
  for (int it = 0; it < n_iter; ++it)
    for (int l = 0; l < ch.n; ++l) {
      const Layer L = ch.L[l];
      const uint32_t idesc = idesc_f16k(256, L.N, L.fmt);
      uint32_t slot = slot0, use = use0;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        slot = slot0; use = use0;
        mbar_wait_cluster(b.a_ready + t, (ar_par >> t) & 1);
        ar_par ^= 1u << t;
        tc_fence_after();
        TL(10, t, l);
        for (int kb = 0; kb < L.nkb; ++kb) {
          if (t == 0) {
            mbar_wait(b.w_full + slot, use & 1);
            mbar_wait_cluster(b.w_peer + slot, use & 1);
            tc_fence_after();
          }
          const int ks_n = L.ksteps - 4 * kb;
          const uint64_t ad = adesc0 + (uint64_t)(t * (A_BYTES >> 4) + kb * (16384 >> 4));
          const uint64_t bd = bdesc0 + (uint64_t)(slot * (SLOT_BYTES >> 4));
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              if (ks < ks_n) mma_bf16_2cta(tmem + t * 256, ad + 2 * ks, bd + 2 * ks, idesc, (kb | ks) != 0);
            if (t == 1) mma_commit_2cta(b.w_empty + slot, 3);
          }
          __syncwarp();
          if (++slot == NSLOT) { slot = 0; ++use; }
        }
        if (elect_one()) mma_commit_2cta(b.acc_full + t, 3);
        __syncwarp();
        TL(11, t, l);
      }
      slot0 = slot; use0 = use;
    }
}

// ---- epilogue-side helpers (t = tile / warp group 0 or 1) ------------------------------------------------
__device__ __forceinline__ void epi_bar(int t) { named_bar_sync(1 + t, EPI_THREADS); }
// all epilogue threads of group t: "tile t's A operand is written and its accumulator drained".
// Optionally the elected thread also saves the first `bytes` of the A tile to global memory AS IS (k-block-major,
// 128B-swizzled "tile layout": tile i of a [rows, 64 nkb] bf16 tensor lives at byte i * nkb * 16384, k-block kb at
// + kb * 16384, row r at + r * 128, 16-byte chunk c at position c ^ (r & 7)) with one TMA bulk store -- coalesced,
// off the epilogue threads' store path, and exactly the layout the weight-gradient kernel consumes.
__device__ __forceinline__ void signal_a_ready(const Bars& b, int t, uint32_t rank, void* gdst = nullptr,
                                               const uint8_t* ssrc = nullptr, uint32_t bytes = 0) {
  tc_fence_before();
  fence_proxy_async();
  epi_bar(t);
  if ((threadIdx.x & (EPI_THREADS - 1)) == 0) {
    TL(4, t, 0);
    if (gdst) { bulk_s2g(gdst, ssrc, bytes); bulk_commit(); }
    if (rank == 0) mbar_arrive_local(b.a_ready + t);
    else mbar_arrive_remote(b.a_ready + t, 0);
  }
}
// before group t overwrites its A tile again: the bulk store issued by signal_a_ready must have read it
__device__ __forceinline__ void drain_store(int t) {
  if ((threadIdx.x & (EPI_THREADS - 1)) == 0) bulk_wait_read0();
  epi_bar(t);
}
__device__ __forceinline__ void wait_acc(const Bars& b, int t, uint32_t& par) {
  mbar_wait(b.acc_full + t, par);
  par ^= 1u;
  tc_fence_after();
}
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// write 16 consecutive columns [c0, c0+16) of this thread's row as fp16 (a FORWARD operand) into an A tile
__device__ __forceinline__ void store_a16(uint8_t* sA, int row, int c0, const float* v) {
  const int kb = c0 >> 6, ch0 = (c0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint4 u;
    u.x = pack_f16(v[8 * q + 0], v[8 * q + 1]); u.y = pack_f16(v[8 * q + 2], v[8 * q + 3]);
    u.z = pack_f16(v[8 * q + 4], v[8 * q + 5]); u.w = pack_f16(v[8 * q + 6], v[8 * q + 7]);
    *reinterpret_cast<uint4*>(sA + kb * 16384 + sw128_off(row, ch0 + q)) = u;
  }
}
__device__ __forceinline__ void store_g16(__nv_bfloat16* dst, const float* v) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint4 u;
    u.x = pack_bf16(v[8 * q + 0], v[8 * q + 1]); u.y = pack_bf16(v[8 * q + 2], v[8 * q + 3]);
    u.z = pack_bf16(v[8 * q + 4], v[8 * q + 5]); u.w = pack_bf16(v[8 * q + 6], v[8 * q + 7]);
    reinterpret_cast<uint4*>(dst)[q] = u;
  }
}

// write 32 consecutive columns [c0, c0+32) of this thread's row as bf16 into an A tile
__device__ __forceinline__ void store_a32(uint8_t* sA, int row, int c0, const float* v) {
  const int kb = c0 >> 6, ch0 = (c0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack_bf16(v[8 * q + 0], v[8 * q + 1]); u.y = pack_bf16(v[8 * q + 2], v[8 * q + 3]);
    u.z = pack_bf16(v[8 * q + 4], v[8 * q + 5]); u.w = pack_bf16(v[8 * q + 6], v[8 * q + 7]);
    *reinterpret_cast<uint4*>(sA + kb * 16384 + sw128_off(row, ch0 + q)) = u;
  }
}
__device__ __forceinline__ void store_g32(__nv_bfloat16* dst, const float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack_bf16(v[8 * q + 0], v[8 * q + 1]); u.y = pack_bf16(v[8 * q + 2], v[8 * q + 3]);
    u.z = pack_bf16(v[8 * q + 4], v[8 * q + 5]); u.w = pack_bf16(v[8 * q + 6], v[8 * q + 7]);
    reinterpret_cast<uint4*>(dst)[q] = u;
  }
}

}  // namespace eng
