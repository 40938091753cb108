// Shared device-side building blocks of the sm_100a kernels (fp32 strict-parity path).
//
// Tile conventions used by every message-passing kernel
//   * a "tile" is TE = 8 entities (edges or nodes) x O = 16 orientations = TM = 128 rows of C = 64
//     channels, staged in shared memory with a padded row stride LDT = 68 floats (272 B) so that the
//     two rows a warp touches per A-operand load sit in different banks;
//   * CTA = 256 threads; thread t owns orientation `o = t >> 4` and channel group `cg = t & 15`
//     (channels 4cg..4cg+3) of EVERY entity of the tile, i.e. rows r_j = 16 j + o, j = 0..7.
//     Because one thread sees all entities of the tile for its (o, channels) slice, segmented sums
//     over edges are plain sequential adds in CSR order: deterministic, no atomics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/grl_b200.h"

namespace grl {

constexpr int kO = GRL_NUM_ORI;    // 16
constexpr int kC = GRL_CHANNELS;   // 64
constexpr int kH = GRL_HIDDEN;     // 256
constexpr int kTE = 8;             // entities per tile
constexpr int kTM = kTE * kO;      // 128 rows per tile
constexpr int kLDT = 68;           // padded smem row stride (floats)
constexpr int kThreads = 256;
constexpr int kRow = kO * kC;      // 1024 floats per latent row
constexpr int kTileFloats = kTM * kLDT;
constexpr int kWFloats = kC * kC;  // one 64x64 weight chunk

// ---- host helpers (grl_util.cu) -------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);
int sm_count();
int ensure_dynamic_smem(const void* func, int bytes);
int make_row_tensor_map(void* out_tensor_map_128B, const float* base, long long n_nodes, int box_rows = 16);  // grl_util.cu

#define GRL_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::grl::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

// ---- math -----------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// GELU and its derivative for the bf16 tensor-core path.  The epilogues of those kernels are instruction-bound
// on the activation (ncu: 75 % of the issued instructions of the node kernels), and their results are rounded
// to bf16 operands (relative step 3.9e-3) right afterwards, so the default is the tanh form evaluated with ONE
// MUFU.TANH:  gelu(x) ~ 0.5 x (1 + tanh(k (x + c x^3))),  |error| <= 4.8e-4 on gelu, 8.7e-4 on gelu' (6 + 5 FP32
// ops).  -DGRL_TC_EXACT_GELU selects the erf form by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, one shared
// exponential, 2 MUFU + ~14 FP32 ops).  The strict fp32 kernels always use erff.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_fast(float x, float& y, float& dy) {
#ifdef GRL_TC_EXACT_GELU
  const float u = __expf(-0.5f * x * x);                                   // exp(-z^2), z = |x| / sqrt(2)
  const float t = __fdividef(1.0f, fmaf(0.23164189f, fabsf(x), 1.0f));     // 1 / (1 + p z), p / sqrt(2) folded in
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float tail = 0.5f * (poly * t) * u;                                // 0.5 * (1 - erf(z))
  const float cdf = x >= 0.f ? 1.0f - tail : tail;
  y = x * cdf;
  dy = fmaf(x * u, 0.39894228040143267794f, cdf);
#else
  constexpr float k = 0.7978845608028654f, kc = 0.7978845608028654f * 0.044715f;
  const float t2 = x * x;
  const float th = tanh_approx(x * fmaf(t2, kc, k));
  const float hx = 0.5f * x;
  y = fmaf(hx, th, hx);
  dy = fmaf(hx * fmaf(-th, th, 1.0f), fmaf(t2, 3.0f * kc, k), fmaf(th, 0.5f, 0.5f));
#endif
}
__device__ __forceinline__ float gelu_fast(float x) {
#ifdef GRL_TC_EXACT_GELU
  float y, dy;
  gelu_fast(x, y, dy);
  return y;
#else
  constexpr float k = 0.7978845608028654f, kc = 0.7978845608028654f * 0.044715f;
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_approx(x * fmaf(x * x, kc, k)), hx);
#endif
}

// ---- cp.async (LDGSTS) ----------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Stage one dense [64][64] fp32 weight chunk (16 KB) into shared memory: 4 x 16 B per thread.
__device__ __forceinline__ void stage_w64(float* __restrict__ dst, const float* __restrict__ src) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int f = threadIdx.x + kThreads * i;  // float4 index 0..1023
    cp_async16(dst + 4 * f, src + 4 * f);
  }
}

// Stage `cnt` (<= 8) CONTIGUOUS latent rows starting at `src` into a padded tile; rows >= cnt are
// zero-filled with plain stores.
__device__ __forceinline__ void stage_rows_contig(float* __restrict__ tile, const float* __restrict__ src, int cnt) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = threadIdx.x + kThreads * i;  // float4 index 0..2047
    const int row = f >> 4, c4 = f & 15;
    float* d = tile + row * kLDT + 4 * c4;
    if ((row >> 4) < cnt) {
      cp_async16(d, src + (size_t)row * kC + 4 * c4);
    } else {
      *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// Gather `cnt` latent rows base[idx[j]] (idx in shared memory) into a padded tile.
__device__ __forceinline__ void stage_rows_gather(float* __restrict__ tile, const float* __restrict__ base,
                                                  const int* __restrict__ idx, int cnt) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = threadIdx.x + kThreads * i;
    const int row = f >> 4, c4 = f & 15;
    const int j = row >> 4;
    float* d = tile + row * kLDT + 4 * c4;
    if (j < cnt) {
      cp_async16(d, base + (size_t)idx[j] * kRow + (row & 15) * kC + 4 * c4);
    } else {
      *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// ---- tile GEMMs (FFMA) ------------------------------------------------------------------------
// acc[j][c] += sum_k A[16 j + o][k] * B[k][4 cg + c]      A: [128][lda] smem, B: [K][64] smem
template <int K>
__device__ __forceinline__ void gemm_tile(const float* __restrict__ As, int lda, const float* __restrict__ Bs, int o,
                                          int cg, float (&acc)[kTE][4]) {
  const float* a0 = As + o * lda;
  const float* b0 = Bs + 4 * cg;
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    const float4 b_0 = *reinterpret_cast<const float4*>(b0 + (k + 0) * kC);
    const float4 b_1 = *reinterpret_cast<const float4*>(b0 + (k + 1) * kC);
    const float4 b_2 = *reinterpret_cast<const float4*>(b0 + (k + 2) * kC);
    const float4 b_3 = *reinterpret_cast<const float4*>(b0 + (k + 3) * kC);
#pragma unroll
    for (int j = 0; j < kTE; ++j) {
      const float4 a = *reinterpret_cast<const float4*>(a0 + j * 16 * lda + k);
      acc[j][0] = fmaf(a.x, b_0.x, acc[j][0]);
      acc[j][1] = fmaf(a.x, b_0.y, acc[j][1]);
      acc[j][2] = fmaf(a.x, b_0.z, acc[j][2]);
      acc[j][3] = fmaf(a.x, b_0.w, acc[j][3]);
      acc[j][0] = fmaf(a.y, b_1.x, acc[j][0]);
      acc[j][1] = fmaf(a.y, b_1.y, acc[j][1]);
      acc[j][2] = fmaf(a.y, b_1.z, acc[j][2]);
      acc[j][3] = fmaf(a.y, b_1.w, acc[j][3]);
      acc[j][0] = fmaf(a.z, b_2.x, acc[j][0]);
      acc[j][1] = fmaf(a.z, b_2.y, acc[j][1]);
      acc[j][2] = fmaf(a.z, b_2.z, acc[j][2]);
      acc[j][3] = fmaf(a.z, b_2.w, acc[j][3]);
      acc[j][0] = fmaf(a.w, b_3.x, acc[j][0]);
      acc[j][1] = fmaf(a.w, b_3.y, acc[j][1]);
      acc[j][2] = fmaf(a.w, b_3.z, acc[j][2]);
      acc[j][3] = fmaf(a.w, b_3.w, acc[j][3]);
    }
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[kTE][4]) {
#pragma unroll
  for (int j = 0; j < kTE; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
}

// Weight-gradient tile: g[i][j] += sum_r A[r][4 ni + i] * B[r][4 mi + j]   (ni = t >> 4, mi = t & 15)
// and, for the lanes with mi == 0, colsum[i] += sum_r A[r][4 ni + i]  (bias gradient, same pass).
template <bool kColSum>
__device__ __forceinline__ void wgrad_tile(const float* __restrict__ As, int lda, const float* __restrict__ Bs, int ldb,
                                           int ni, int mi, float (&g)[4][4], float (&colsum)[4]) {
  const float* a0 = As + 4 * ni;
  const float* b0 = Bs + 4 * mi;
#pragma unroll 4
  for (int r = 0; r < kTM; ++r) {
    const float4 a = *reinterpret_cast<const float4*>(a0 + r * lda);
    const float4 b = *reinterpret_cast<const float4*>(b0 + r * ldb);
    g[0][0] = fmaf(a.x, b.x, g[0][0]); g[0][1] = fmaf(a.x, b.y, g[0][1]);
    g[0][2] = fmaf(a.x, b.z, g[0][2]); g[0][3] = fmaf(a.x, b.w, g[0][3]);
    g[1][0] = fmaf(a.y, b.x, g[1][0]); g[1][1] = fmaf(a.y, b.y, g[1][1]);
    g[1][2] = fmaf(a.y, b.z, g[1][2]); g[1][3] = fmaf(a.y, b.w, g[1][3]);
    g[2][0] = fmaf(a.z, b.x, g[2][0]); g[2][1] = fmaf(a.z, b.y, g[2][1]);
    g[2][2] = fmaf(a.z, b.z, g[2][2]); g[2][3] = fmaf(a.z, b.w, g[2][3]);
    g[3][0] = fmaf(a.w, b.x, g[3][0]); g[3][1] = fmaf(a.w, b.y, g[3][1]);
    g[3][2] = fmaf(a.w, b.z, g[3][2]); g[3][3] = fmaf(a.w, b.w, g[3][3]);
    if (kColSum && mi == 0) {
      colsum[0] += a.x; colsum[1] += a.y; colsum[2] += a.z; colsum[3] += a.w;
    }
  }
}

// Sum over the 16 lanes that share a row (the 16 channel groups of a half-warp).
__device__ __forceinline__ float row_sum16(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  return v;
}

// ---- packed fp32 pairs (sm_100 FFMA2: two fused multiply-adds per issue slot of the fma pipe, each lane rounded exactly like fmaf)
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace grl
