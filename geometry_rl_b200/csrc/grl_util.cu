// Error plumbing, device query and the deterministic cross-CTA partial reduction.
#include <cstdarg>
#include <cstdio>
#include <atomic>
#include <mutex>
#include <utility>
#include <vector>

#include <cuda.h>

#include "grl_common.cuh"

namespace grl {

// cuTensorMapEncodeTiled through the runtime's driver-entry-point query: the library links no libcuda symbol.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// Tensor map over a latent tensor viewed as [n_nodes * 16 orientation rows][64 fp32]: box = 16 rows x 32 channels
// (one node, one channel half = 2 KB), SWIZZLE_128B: the 16-byte chunk c of row r lands at chunk position c ^ (r & 7)
// of its 128-byte line, so threads that own one ROW each (TMEM lane = row) read any chunk conflict-free.
int make_row_tensor_map(void* out, const float* base, long long n_nodes, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return GRL_ECUDA;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)kC, (cuuint64_t)n_nodes * kO};
  const cuuint64_t gstride[1] = {(cuuint64_t)kC * sizeof(float)};
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};  // 16 = one node (gathers), 128 = a contiguous 8-node tile
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim,
                        gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (base %p, %lld nodes)", (int)r, (const void*)base, n_nodes);
    return GRL_ECUDA;
  }
  return GRL_OK;
}

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return GRL_ECUDA;
  }
  return GRL_OK;
}

// SMs the persistent kernels leave free (grl_reserve_sms): a process-wide launch policy, 0 by default.
static std::atomic<int> g_reserved_sms{0};

int sm_count() {
  // immutable per-device capability cache
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148 - g_reserved_sms.load(std::memory_order_relaxed);
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  const int usable = cached[dev] - g_reserved_sms.load(std::memory_order_relaxed);
  return usable > 1 ? usable : 1;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE property of a kernel: set it once per (kernel, device)
// pair.  (A process-wide "already set" latch breaks the second device a process drives.)  The table only ever grows
// and an entry never changes: idempotent capability state, like sm_count()'s cache.
int ensure_dynamic_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::vector<std::pair<const void*, int>> done;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("cudaGetDevice failed");
    return GRL_ECUDA;
  }
  std::lock_guard<std::mutex> lock(mu);
  for (const auto& e : done)
    if (e.first == func && e.second == dev) return GRL_OK;
  const cudaError_t err = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (err != cudaSuccess) {
    set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize=%d): %s", bytes, cudaGetErrorString(err));
    return GRL_ECUDA;
  }
  done.emplace_back(func, dev);
  return GRL_OK;
}

// out[i] (+)= sum_p partials[p][i] in a FIXED order: the partials are split into 8 contiguous segments, each summed
// sequentially by one thread (4 independent running sums, p = 4 q + r), and the 8 segment sums are added in segment
// order.  A block covers 32 outputs x 8 segments, so 32 consecutive threads read 128 contiguous bytes and a
// 296-partial reduction is 10 dependent loads deep instead of 296 (ncu r01b: 40 us for 4096 outputs before).
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int n_partials, int64_t n,
                                                              float* __restrict__ out, int accumulate) {
  __shared__ float seg_sum[8][32];
  const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int per = (n_partials + 7) / 8;
  const int p0 = seg * per, p1 = min(n_partials, p0 + per);
  for (int64_t base = (int64_t)blockIdx.x * 32; base < n; base += (int64_t)gridDim.x * 32) {
    const int64_t i = base + lane;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (i < n) {
      int p = p0;
      for (; p + 3 < p1; p += 4) {
        a0 += partials[(int64_t)p * n + i];
        a1 += partials[(int64_t)(p + 1) * n + i];
        a2 += partials[(int64_t)(p + 2) * n + i];
        a3 += partials[(int64_t)(p + 3) * n + i];
      }
      for (; p < p1; ++p) a0 += partials[(int64_t)p * n + i];
    }
    seg_sum[seg][lane] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (seg == 0 && i < n) {
      float s = seg_sum[0][lane];
#pragma unroll
      for (int k = 1; k < 8; ++k) s += seg_sum[k][lane];
      out[i] = accumulate ? out[i] + s : s;
    }
    __syncthreads();
  }
}

// max |x| over a tensor as the bit pattern of a non-negative float (order-independent -> deterministic);
// NaNs are ignored, the result buffer must be zero on entry.
__global__ void __launch_bounds__(256) absmax_kernel(const float4* __restrict__ x, int64_t n4, const float* __restrict__ tail,
                                                     int n_tail, unsigned int* __restrict__ out_bits) {
  unsigned int m = 0u;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    const float a = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));  // fmaxf drops NaNs
    m = max(m, __float_as_uint(a));
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) m = max(m, __float_as_uint(fabsf(tail[threadIdx.x])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ unsigned int wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = max(m, wm[w]);
    atomicMax(out_bits, m);
  }
}

}  // namespace grl

extern "C" {

int grl_abi_version(void) { return 1; }
const char* grl_last_error(void) { return grl::g_err; }
int grl_sm_count(void) { return grl::sm_count(); }
int grl_reserve_sms(int n) {
  GRL_REQUIRE(n >= 0 && n <= 64, GRL_EINVAL, "grl_reserve_sms: n=%d outside [0, 64]", n);
  grl::g_reserved_sms.store(n, std::memory_order_relaxed);
  return GRL_OK;
}

int grl_reduce_partials(const float* partials, int n_partials, int64_t n_floats, float* out, int accumulate,
                        grl_stream_t stream) {
  GRL_REQUIRE(partials && out && n_partials > 0 && n_floats > 0, GRL_EINVAL, "grl_reduce_partials: bad arguments");
  int64_t blocks = (n_floats + 31) / 32;
  if (blocks > 16 * grl::sm_count()) blocks = 16 * grl::sm_count();
  grl::reduce_partials_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(partials, n_partials, n_floats, out, accumulate);
  return grl::check_launch("grl_reduce_partials");
}

int grl_absmax(const float* x, int64_t n, uint32_t* out_bits, grl_stream_t stream) {
  GRL_REQUIRE(x && out_bits && n > 0, GRL_EINVAL, "grl_absmax: bad arguments");
  GRL_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, GRL_EINVAL, "grl_absmax: x must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(out_bits, 0, sizeof(uint32_t), s) != cudaSuccess) return grl::check_launch("grl_absmax (memset)");
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 8 * grl::sm_count()) blocks = 8 * grl::sm_count();
  if (blocks < 1) blocks = 1;
  grl::absmax_kernel<<<(int)blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(x), n4, x + 4 * n4, (int)(n - 4 * n4), out_bits);
  return grl::check_launch("grl_absmax");
}

}  // extern "C"
