// mlp_tc.cu -- bf16 tensor-core mode of the per-pair field kernels (tcgen05.mma + TMEM, weights by bulk copy).
// (work in progress: first the single-GEMM building-block test entry point)
#include "common.cuh"
#include "umma.cuh"

using namespace tc;

// -------------------------------------------------------------------------------------------------
// building-block self test: out[128][N] = A[128][K] (bf16 row-major, K multiple of 16, <= 256) @ Wp^T, where Wp is
// the packed (k-block major, SW128) image of W [N][K].  Exercises: generic-proxy writes of the A tile in the
// swizzled layout, bulk-copy weight load on an mbarrier, tcgen05.mma issue + commit, tcgen05.ld epilogue.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
k_tc_gemm_test(const __nv_bfloat16* __restrict__ A, const uint8_t* __restrict__ Wp, int N, int K, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                 // 4 k-blocks x 16 KB
  uint8_t* sW = smem + 65536;         // 4 k-blocks x N*128 B
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + 65536 + 131072);
  uint64_t* bar_m = bar_w + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_m + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = (K + 63) / 64;
  if (tid == 0) { mbar_init(bar_w, 1); mbar_init(bar_m, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(s_tmem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)nkb * N * 128;
    mbar_expect_tx(bar_w, bytes);
    for (int kb = 0; kb < nkb; ++kb) bulk_g2s(sW + (size_t)kb * N * 128, Wp + (size_t)kb * N * 128, N * 128, bar_w);
  }
  // A tile: thread -> (row, chunks)
  for (int e = tid; e < 128 * nkb * 8; e += 256) {
    int row = e / (nkb * 8), cc = e % (nkb * 8), kb = cc >> 3, ch = cc & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    int k0 = kb * 64 + ch * 8;
    if (k0 < K) v = *reinterpret_cast<const uint4*>(A + (size_t)row * K + k0);
    *reinterpret_cast<uint4*>(sA + kb * 16384 + sw128_off(row, ch)) = v;
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    mbar_wait(bar_w, 0);
    tc_fence_after();
    const uint32_t idesc = idesc_bf16(128, N);
    int acc = 0;
    for (int k = 0; k < K; k += 16) {
      int kb = k >> 6, ks = (k & 63) >> 4;
      uint64_t ad = smem_desc_sw128(smem_u32(sA) + kb * 16384 + ks * 32);
      uint64_t bd = smem_desc_sw128(smem_u32(sW) + kb * N * 128 + ks * 32);
      mma_bf16(tmem, ad, bd, idesc, acc);
      acc = 1;
    }
    mma_commit(bar_m);
  }
  mbar_wait(bar_m, 0);
  tc_fence_after();
  const int row = 32 * (warp & 3) + lane;
  const int half = warp >> 2;
  for (int c0 = half * 128; c0 < half * 128 + 128 && c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) out[(size_t)row * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

extern "C" int spf_tc_gemm_test(const void* A, const void* Wp, int32_t N, int32_t K, float* out, void* stream_) {
  if (!A || !Wp || !out) return SPF_ERR_INVALID;
  if (N % 16 || N < 16 || N > 256 || K % 16 || K < 16 || K > 256) return SPF_ERR_UNSUPPORTED;
  const int smem = 65536 + 131072 + 1024;
  SPF_CUDA(cudaFuncSetAttribute(k_tc_gemm_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "tc test attr");
  k_tc_gemm_test<<<1, 256, smem, (cudaStream_t)stream_>>>((const __nv_bfloat16*)A, (const uint8_t*)Wp, N, K, out);
  SPF_CHECK_LAUNCH("k_tc_gemm_test");
  return SPF_OK;
}
