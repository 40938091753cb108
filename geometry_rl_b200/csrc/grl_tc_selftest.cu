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

// ---- TS-mode self-test: D[128 x 64] = fp16(A[128 x K]) * fp16(B[64 x K])^T with A fed from TENSOR MEMORY ------------
namespace grl {
template <int K>
__global__ void __launch_bounds__(128) tc_selftest_ts_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                             float* __restrict__ D) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __half* sB = reinterpret_cast<__half*>(smem_raw);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 256);
  tc::stage_weight_f16(sB, B, 64, K, K);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base, lane_addr = tmem + ((uint32_t)(32 * warp) << 16);
  constexpr uint32_t kColA = 128;  // A: K / 2 packed columns from column 128; D: columns 0..63
  // thread = row: pack its K values into fp16 pairs and store them into its TMEM lane, 16 columns (32 values) at a time
#pragma unroll 1
  for (int c = 0; c < K / 2; c += 16) {
    uint32_t r[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const __half2 h = __floats2half2_rn(A[(size_t)tid * K + 2 * (c + j)], A[(size_t)tid * K + 2 * (c + j) + 1]);
      r[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    tc::tmem_st16(lane_addr + kColA + c, r);
  }
  tc::tmem_st_wait();
  tc::tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc::tc_fence_after();
    tc::issue_mma_ts(tmem, tmem + kColA, tc::view_k(tc::smem_u32(sB), 64), tc::idesc_f16_ex(128, 64, 0, 0, 0, 0), K / 16, false);
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
#pragma unroll 1
  for (int c0 = 0; c0 < 64; c0 += 16) {
    float v[16];
    tc::tmem_ld16(lane_addr + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)tid * 64 + c0 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}
}  // namespace grl

extern "C" int grl_tc_selftest_gemm_ts(const float* A, const float* B, float* D, int K, grl_stream_t stream) {
  GRL_REQUIRE(A && B && D, GRL_EINVAL, "grl_tc_selftest_gemm_ts: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if (K == 64) grl::tc_selftest_ts_kernel<64><<<1, 128, 64 * 64 * 2, s>>>(A, B, D);
  else if (K == 256) grl::tc_selftest_ts_kernel<256><<<1, 128, 64 * 256 * 2, s>>>(A, B, D);
  else GRL_REQUIRE(false, GRL_EUNSUPPORTED, "grl_tc_selftest_gemm_ts: K=%d not instantiated", K);
  return grl::check_launch("grl_tc_selftest_gemm_ts");
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

// ---- latency probe: the hand-off costs that bound the warp-specialised kernels -------------------------------------
// One CTA, 3 warps: warp 0 = "MMA warp", warp 1 = "epilogue warp", warp 2 idle.  Every figure is SM clock cycles
// (clock64), median-free: the probe repeats each measurement `kRep` times and reports the minimum and the mean.
//   out[2 i], out[2 i + 1] = min, mean of measurement i:
//   0  n = 1 MMA (M128 N64 K16) issue + commit -> mbarrier wait returns in the ISSUING thread
//   1  n = 4          2  n = 8          3  n = 16
//   4  fence.proxy.async after two 16-byte shared stores
//   5  mbarrier.arrive in warp 1 -> try_wait returns in warp 0 (one way)
//   6  tcgen05.ld 32x32b.x16 + wait::ld
//   7  round trip measured in warp 1: arrive -> warp 0 wakes, issues 4 MMAs + commit -> warp 1's wait returns
//   8  the same with the st.shared + fence.proxy.async + tcgen05 fences of a real epilogue in front of the arrive
namespace grl {
constexpr int kProbeRep = 64;
__global__ void __launch_bounds__(96) tc_latency_probe_kernel(long long* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_mma, bar_a, bar_b;
  __shared__ uint32_t tmem_base;
  __shared__ long long t_arrive[kProbeRep];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 2 * 32768 / 16; i += 96) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    tc::mbar_init(&bar_mma, 1);
    tc::mbar_init(&bar_a, 1);
    tc::mbar_init(&bar_b, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 128);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base, sa = tc::smem_u32(smem_raw), sb = sa + 32768;
  const tc::OpView va = tc::view_k(sa, 128), vb = tc::view_k(sb, 64);
  constexpr uint32_t idesc = tc::idesc_bf16(128, 64);
  uint32_t par = 0;
  auto record = [&](int slot, long long mn, long long sum) {
    out[2 * slot] = mn;
    out[2 * slot + 1] = sum / kProbeRep;
  };
  // ---- 0..3: MMA + commit latency seen by the issuing thread
  if (tid == 0) {
    const int ns[4] = {1, 4, 8, 16};
    for (int c = 0; c < 4; ++c) {
      long long mn = 1ll << 60, sum = 0;
      for (int r = 0; r < kProbeRep; ++r) {
        const long long t0 = clock64();
        for (int k = 0; k < ns[c]; ++k) tc::issue_mma(tmem, va, vb, idesc, 1, k > 0);
        tc::mma_commit(&bar_mma);
        tc::mbar_wait(&bar_mma, par);
        const long long t1 = clock64();
        par ^= 1u;
        mn = min(mn, t1 - t0);
        sum += t1 - t0;
      }
      record(c, mn, sum);
    }
  }
  // ---- 9..13: 16 back-to-back (unrolled) MMAs + commit -> wait for the operand views the kernels use
  //   9  A K-major, B K-major, N = 64      10  A K-major, B MN-major, N = 64     11  A MN-major, B MN-major, N = 64
  //   12 A MN-major, B MN-major, N = 80    13  A K-major, B K-major, N = 128 (8 MMAs)
  if (tid == 0) {
    for (int c = 0; c < 5; ++c) {
      long long mn = 1ll << 60, sum = 0;
      for (int r = 0; r < kProbeRep; ++r) {
        const long long t0 = clock64();
        if (c == 0) {
          tc::issue_mma(tmem, tc::view_k(sa, 128), tc::view_k(sb, 64), tc::idesc_f16_ex(128, 64, 0, 0, 1, 1), 4, false);
          tc::issue_mma(tmem, tc::view_k(sa, 128), tc::view_k(sb, 64), tc::idesc_f16_ex(128, 64, 0, 0, 1, 1), 4, true);
          tc::issue_mma(tmem, tc::view_k(sa, 128), tc::view_k(sb, 64), tc::idesc_f16_ex(128, 64, 0, 0, 1, 1), 4, true);
          tc::issue_mma(tmem, tc::view_k(sa, 128), tc::view_k(sb, 64), tc::idesc_f16_ex(128, 64, 0, 0, 1, 1), 4, true);
        } else if (c == 1) {
          tc::issue_mma(tmem, tc::view_k(sa, 128), tc::view_mn(sb, 64), tc::idesc_f16_ex(128, 64, 0, 1, 1, 1), 4, false);
          tc::issue_mma(tmem, tc::view_k(sa, 128), tc::view_mn(sb, 64), tc::idesc_f16_ex(128, 64, 0, 1, 1, 1), 4, true);
          tc::issue_mma(tmem, tc::view_k(sa, 128), tc::view_mn(sb, 64), tc::idesc_f16_ex(128, 64, 0, 1, 1, 1), 4, true);
          tc::issue_mma(tmem, tc::view_k(sa, 128), tc::view_mn(sb, 64), tc::idesc_f16_ex(128, 64, 0, 1, 1, 1), 4, true);
        } else if (c == 2) {
          tc::issue_mma(tmem, tc::view_mn(sa, 128), tc::view_mn(sb, 128), tc::idesc_f16_ex(128, 64, 1, 1, 1, 1), 8, false);
          tc::issue_mma(tmem, tc::view_mn(sa, 128), tc::view_mn(sb, 128), tc::idesc_f16_ex(128, 64, 1, 1, 1, 1), 8, true);
        } else if (c == 3) {
          tc::issue_mma(tmem, tc::view_mn(sa, 128), tc::view_mn(sb, 128), tc::idesc_f16_ex(128, 80, 1, 1, 1, 1), 8, false);
          tc::issue_mma(tmem, tc::view_mn(sa, 128), tc::view_mn(sb, 128), tc::idesc_f16_ex(128, 80, 1, 1, 1, 1), 8, true);
        } else {
          tc::issue_mma(tmem, tc::view_k(sa, 128), tc::view_k(sb, 128), tc::idesc_f16_ex(128, 128, 0, 0, 1, 1), 4, false);
          tc::issue_mma(tmem, tc::view_k(sa, 128), tc::view_k(sb, 128), tc::idesc_f16_ex(128, 128, 0, 0, 1, 1), 4, true);
        }
        tc::mma_commit(&bar_mma);
        tc::mbar_wait(&bar_mma, par);
        const long long t1 = clock64();
        par ^= 1u;
        mn = min(mn, t1 - t0);
        sum += t1 - t0;
      }
      record(9 + c, mn, sum);
    }
  }
  __syncthreads();
  // ---- 4: fence.proxy.async after shared stores;  6: tcgen05.ld
  if (warp == 1) {
    long long mn = 1ll << 60, sum = 0;
    for (int r = 0; r < kProbeRep; ++r) {
      const long long t0 = clock64();
      *reinterpret_cast<uint4*>(smem_raw + tid * 16) = make_uint4(r, r, r, r);
      *reinterpret_cast<uint4*>(smem_raw + 4096 + tid * 16) = make_uint4(r, r, r, r);
      tc::fence_async_smem();
      const long long t1 = clock64();
      mn = min(mn, t1 - t0);
      sum += t1 - t0;
    }
    if (lane == 0) record(4, mn, sum);
    mn = 1ll << 60, sum = 0;
    float acc = 0.f;
    for (int r = 0; r < kProbeRep; ++r) {
      const long long t0 = clock64();
      float v[16];
      tc::tmem_ld16(tmem + ((uint32_t)32 << 16), v);
      acc += v[r & 15];
      const long long t1 = clock64();
      mn = min(mn, t1 - t0);
      sum += t1 - t0;
    }
    if (lane == 0) {
      record(6, mn, sum);
      if (acc == 123.456f) out[63] = 1;
    }
  }
  __syncthreads();
  // ---- 5: one-way arrive -> wait;  7 / 8: round trips
  for (int mode = 0; mode < 3; ++mode) {
    uint32_t pa = 0, pb = 0;
    long long mn = 1ll << 60, sum = 0;
    for (int r = 0; r < kProbeRep; ++r) {
      __syncthreads();
      if (warp == 1) {
        // desynchronise a little so the waiter is already parked in try_wait
        for (int spin = 0; spin < 200; ++spin) asm volatile("nanosleep.u32 20;");
        const long long t0 = clock64();
        if (mode == 2) {
          *reinterpret_cast<uint4*>(smem_raw + tid * 16) = make_uint4(r, r, r, r);
          *reinterpret_cast<uint4*>(smem_raw + 4096 + tid * 16) = make_uint4(r, r, r, r);
          tc::fence_async_smem();
          tc::tc_fence_before();
          __syncwarp();
        }
        if (lane == 0) {
          t_arrive[r] = t0;
          tc::mbar_arrive(&bar_a);
        }
        if (mode >= 1) {
          tc::mbar_wait(&bar_b, pb);
          pb ^= 1u;
          tc::tc_fence_after();
          const long long t1 = clock64();
          mn = min(mn, t1 - t0);
          sum += t1 - t0;
        }
      } else if (warp == 0) {
        tc::mbar_wait(&bar_a, pa);
        pa ^= 1u;
        const long long t1 = clock64();
        if (mode == 0) {
          __threadfence_block();
          const long long dt = t1 - t_arrive[r];
          mn = min(mn, dt);
          sum += dt;
        } else {
          tc::tc_fence_after();
          if (tc::elect_one()) {
            tc::issue_mma(tmem, va, vb, idesc, 4, false);
            tc::mma_commit(&bar_b);
          }
          __syncwarp();
        }
      }
    }
    __syncthreads();
    if (mode == 0 && tid == 0) record(5, mn, sum);
    if (mode == 1 && tid == 32) record(7, mn, sum);
    if (mode == 2 && tid == 32) record(8, mn, sum);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}
}  // namespace grl

extern "C" int grl_tc_latency_probe(long long* out64, grl_stream_t stream) {
  GRL_REQUIRE(out64, GRL_EINVAL, "grl_tc_latency_probe: null pointer");
  const int smem = 2 * 32768;
  if (grl::ensure_dynamic_smem((const void*)grl::tc_latency_probe_kernel, smem) != GRL_OK) return GRL_ECUDA;
  grl::tc_latency_probe_kernel<<<1, 96, smem, (cudaStream_t)stream>>>(out64);
  return grl::check_launch("grl_tc_latency_probe");
}
