// Self-test of the tcgen05 building blocks (grl_tc.cuh): D[128 x N] = bf16(A[128 x K]) * bf16(B[N x K])^T with
// fp32 accumulation in TMEM.  One CTA, 128 threads.  Used by tests/test_gpu_tc.py to pin the shared-memory
// descriptor / instruction descriptor / TMEM lane mapping before the fused kernels rely on them.
#include "grl_common.cuh"
#include "grl_tc.cuh"

namespace grl {

template <int N, int K>
__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          float* __restrict__ D) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sB = sA + 128 * K;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base, N < 32 ? 32 : N);
  tc::stage_weight_bf16(sA, A, 128, K, K);
  tc::stage_weight_bf16(sB, B, N, K, K);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    tc::issue_gemm<N, K>(tmem, tc::smem_u32(sA), 128, tc::smem_u32(sB), N, false);
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
#pragma unroll 1
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)tid * N + c0 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, N < 32 ? 32 : N);
}

template <int N, int K>
static int launch_selftest(const float* A, const float* B, float* D, cudaStream_t s) {
  const int smem = (128 + N) * K * 2;
  cudaFuncSetAttribute(tc_selftest_kernel<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  tc_selftest_kernel<N, K><<<1, 128, smem, s>>>(A, B, D);
  return check_launch("grl_tc_selftest_gemm");
}

}  // namespace grl

extern "C" int grl_tc_selftest_gemm(const float* A, const float* B, float* D, int N, int K, grl_stream_t stream) {
  GRL_REQUIRE(A && B && D, GRL_EINVAL, "grl_tc_selftest_gemm: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if (N == 64 && K == 16) return grl::launch_selftest<64, 16>(A, B, D, s);
  if (N == 64 && K == 64) return grl::launch_selftest<64, 64>(A, B, D, s);
  if (N == 256 && K == 64) return grl::launch_selftest<256, 64>(A, B, D, s);
  if (N == 64 && K == 256) return grl::launch_selftest<64, 256>(A, B, D, s);
  GRL_REQUIRE(false, GRL_EUNSUPPORTED, "grl_tc_selftest_gemm: (N,K)=(%d,%d) not instantiated", N, K);
}

// ---- raw debug hook: caller supplies the shared-memory images and every descriptor field ------------------
namespace grl {
__global__ void __launch_bounds__(128) tc_debug_kernel(const uint4* __restrict__ a_img, int a_bytes, const uint4* __restrict__ b_img,
                                                       int b_bytes, float* __restrict__ D, int N, int n_ksteps,
                                                       uint32_t lbo_a, uint32_t sbo_a, uint32_t adv_a, uint32_t lbo_b,
                                                       uint32_t sbo_b, uint32_t adv_b, uint32_t idesc, uint32_t desc_hi_bits) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sA = smem_raw;
  unsigned char* sB = smem_raw + ((a_bytes + 127) / 128) * 128;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 256);
  for (int i = tid; i < a_bytes / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = a_img[i];
  for (int i = tid; i < b_bytes / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = b_img[i];
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    for (int k = 0; k < n_ksteps; ++k) {
      uint64_t da = tc::smem_desc(tc::smem_u32(sA) + k * adv_a, lbo_a, sbo_a) | ((uint64_t)desc_hi_bits << 32);
      uint64_t db = tc::smem_desc(tc::smem_u32(sB) + k * adv_b, lbo_b, sbo_b) | ((uint64_t)desc_hi_bits << 32);
      tc::mma_bf16(tmem, da, db, idesc, k > 0 ? 1u : 0u);
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 16; ++i) D[(size_t)tid * N + c0 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}
}  // namespace grl

extern "C" int grl_tc_debug_mma(const void* a_img, int a_bytes, const void* b_img, int b_bytes, float* D, int N, int n_ksteps,
                                uint32_t lbo_a, uint32_t sbo_a, uint32_t adv_a, uint32_t lbo_b, uint32_t sbo_b, uint32_t adv_b,
                                uint32_t idesc, uint32_t desc_hi_bits, grl_stream_t stream) {
  GRL_REQUIRE(a_img && b_img && D && N > 0 && N <= 256 && a_bytes % 16 == 0 && b_bytes % 16 == 0, GRL_EINVAL,
              "grl_tc_debug_mma: bad arguments");
  const int smem = ((a_bytes + 127) / 128) * 128 + b_bytes;
  GRL_REQUIRE(smem <= 200 * 1024, GRL_EUNSUPPORTED, "grl_tc_debug_mma: images too large");
  cudaFuncSetAttribute(grl::tc_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  grl::tc_debug_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const uint4*)a_img, a_bytes, (const uint4*)b_img, b_bytes, D, N,
                                                              n_ksteps, lbo_a, sbo_a, adv_a, lbo_b, sbo_b, adv_b, idesc,
                                                              desc_hi_bits);
  return grl::check_launch("grl_tc_debug_mma");
}
