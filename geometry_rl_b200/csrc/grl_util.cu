// Error plumbing, device query and the deterministic cross-CTA partial reduction.
#include <cstdarg>
#include <cstdio>

#include "grl_common.cuh"

namespace grl {

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

int sm_count() {
  // immutable per-device capability cache (the only global state behind the ABI)
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// out[i] (+)= sum_p partials[p][i] in fixed p order.
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int n_partials, int64_t n, float* __restrict__ out,
                                       int accumulate) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < n_partials; ++p) s += partials[(int64_t)p * n + i];
    out[i] = accumulate ? out[i] + s : s;
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

int grl_reduce_partials(const float* partials, int n_partials, int64_t n_floats, float* out, int accumulate,
                        grl_stream_t stream) {
  GRL_REQUIRE(partials && out && n_partials > 0 && n_floats > 0, GRL_EINVAL, "grl_reduce_partials: bad arguments");
  const int threads = 256;
  int64_t blocks = (n_floats + threads - 1) / threads;
  if (blocks > 4 * grl::sm_count()) blocks = 4 * grl::sm_count();
  grl::reduce_partials_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(partials, n_partials, n_floats, out,
                                                                                  accumulate);
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
