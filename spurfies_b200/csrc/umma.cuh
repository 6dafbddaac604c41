// umma.cuh -- inline-PTX wrappers for the Blackwell (sm_100a) tensor path used by mlp_tc.cu:
// tcgen05.mma (bf16 x bf16 -> fp32 in TMEM), TMEM alloc / tcgen05.ld, mbarrier, bulk-copy (TMA engine) loads.
// Descriptor encodings follow the PTX ISA tables (cross-checked against the CuTe reference structs
// cute/arch/mma_sm100_desc.hpp: SmemDescriptor / InstrDescriptor).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---- proxies / fences -------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- bulk copy global -> shared (TMA engine, SASS UBLKCP), completes on an mbarrier -------------
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- bulk copy shared -> global (TMA engine), tracked by bulk async-groups of the issuing thread ------
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---- TMEM -------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane (lane = 32*(warp%4) + laneid)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------------------------
// K-major operand tile stored as 128-byte rows (64 bf16), 8-row groups of 1024 B, 128B swizzle
// (16-byte chunk index XOR (row & 7)).  SBO = 1024 B, LBO field = 1 (ignored for swizzled K-major),
// descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.  Tile base must be 1024-B aligned;
// K is advanced inside the swizzle atom by adding k*32 B to the start address.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16, A = B = bf16 (format 1), D = fp32 (format 1), both K-major, dense, no negate
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Operand formats of kind::f16 (instruction-descriptor bits 7..9 = A, 10..12 = B: 0 = fp16, 1 = bf16).  FORWARD operands
// (inputs, activations, weights) are fp16: 11 significant bits instead of bf16's 8, at the same tensor throughput, so 8x
// fewer LeakyReLU pre-activations land on the wrong side of zero than with bf16 operands (each such flip moves that
// unit's whole gradient contribution: tools/bf16_grad_study.py).  GRADIENT operands (dZ, the d sdf / d input chain) stay
// bf16: they span many decades and are not range-safe in fp16.  `fmt` bit 0 / bit 1 = A / B operand is bf16.
constexpr int FMT_F16 = 0, FMT_A_BF16 = 1, FMT_B_BF16 = 2, FMT_BF16 = 3;
__host__ __device__ constexpr uint32_t idesc_f16k(int M, int N, int fmt) {
  return (1u << 4) | ((uint32_t)(fmt & 1) << 7) | ((uint32_t)((fmt >> 1) & 1) << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_f16k_mn(int M, int N, int fmt) {
  return idesc_f16k(M, N, fmt) | (1u << 15) | (1u << 16);
}
// D[tmem] (+)= A[smem] * B[smem]^T, one K=16 step; issued by ONE thread
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// MN-major operand tile (the contraction index k is the SLOW index in memory): 64-element (128 B) rows along M/N,
// 8 k-rows per 1024-B swizzle atom; panels of 64 M/N elements are `panel_bytes` apart (LBO), 8-row k groups 1024 B
// apart (SBO).  One K=16 MMA step consumes two k groups: advance the start address by 2048 B per step.
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr, uint32_t panel_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((panel_bytes >> 4) & 0x3FFFu) << 16) | (64ull << 32) | (1ull << 46) |
         (2ull << 61);
}
// as idesc_bf16 but both operands MN-major (transposed in memory)
__host__ __device__ constexpr uint32_t idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


// ---- CTA pair (cta_group::2) / cluster helpers ----------------------------------------------------
// Two CTAs of a cluster (one TPC) run ONE tcgen05.mma of M = 256: each CTA supplies its own 128 rows of A and its own
// N/2 rows of B from the SAME shared-memory offsets, and receives its 128 rows x N columns of D in its own TMEM.
// Only the leader (cluster rank 0) issues; completion is multicast to the barriers at the same offset in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
// wait on a barrier other CTAs of the cluster arrive on.  Default (cta-scope) acquire, as CUTLASS's ClusterBarrier::wait:
// a cluster-scope acquire makes ptxas emit CCTL.IVALL (whole-L1 invalidate) after every wait, and the waiter (the MMA
// issuer) consumes the guarded shared memory through the async proxy only.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WAIT_DONE_C;\n\t"
      "bra WAIT_LOOP_C;\n\t"
      "WAIT_DONE_C:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// Arrivals use the default .release.cta semantics (what CUTLASS's ClusterBarrier::arrive emits): a cluster-scope release
// compiles to MEMBAR.ALL.GPU + CCTL on the epilogue -> MMA critical path.  The data these barriers order is shared memory
// written by the arriving CTA's own threads and already made visible to the async proxy by fence.proxy.async + bar.sync.
__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {  // warp 0 of BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (when all MMAs issued so far by this thread have completed) on `bar` in every CTA of `mask`
__device__ __forceinline__ void mma_commit_2cta(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// one (compiler-visible) elected lane of a fully active warp: lets ptxas issue uniform-datapath instructions
// (tcgen05.mma / commit) without the per-thread serialisation loop it wraps around them in divergent code
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
// named barrier among `nthreads` threads (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// byte offset of (row, 16-byte chunk) inside one [rows x 64 bf16] SW128 k-block
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) { return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// two fp32 -> packed fp16 (a in the low half), round to nearest.  (No .satfinite: ptxas accepts it for sm_100a and emits
// F2FP.SATFINITE.F16.F32, which the B200 rejects at run time as an illegal instruction.)  |x| > 65504 becomes inf like in
// any fp16 forward; the MLP inputs are bounded (latents, |x - p| <= 0.05, sin / cos) and the weight images are clamped
// when they are packed.
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// LeakyReLU on a packed fp16 pair: max(h, slope * h) as two f16x2 instructions for two elements (the fp32 form costs two
// per element).  Applied AFTER the fp32 bias add and the rounding to fp16, so the sign of every pre-activation -- the
// only thing the backward's masks depend on -- is exactly the fp32 one; the negative branch is rounded twice (2^-12
// relative of a value that is already 100x smaller).  slope2 = the slope as a packed fp16 pair (0.01 -> 0x211F211F).
__device__ __forceinline__ uint32_t leaky_f16x2(uint32_t h, uint32_t slope2) {
  uint32_t m, r;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(m) : "r"(h), "r"(slope2));
  asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(h), "r"(m));
  return r;
}
constexpr uint32_t LEAKY_H2 = 0x211F211Fu;   // 0.01 (0.010002) as fp16 x 2
constexpr uint32_t ONE_H2 = 0x3C003C00u;     // 1.0 as fp16 x 2 (a linear layer through the same code path)
__device__ __forceinline__ float f16_round(float a) {   // the value pack_f16 stores, back in fp32
  unsigned short h;
  asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(a));
  float r;
  asm("cvt.f32.f16 %0, %1;" : "=f"(r) : "h"(h));
  return r;
}

}  // namespace tc
