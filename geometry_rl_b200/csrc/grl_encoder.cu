// K5: one post-LN transformer encoder layer over the tokens of a graph, forward and backward, fp32.
//   a  = MHA(x) = concat_h softmax((x Wq_h^T + bq_h)(x Wk_h^T + bk_h)^T / sqrt(32)) (x Wv_h^T + bv_h) Wo^T + bo
//   x1 = LayerNorm1(x + a)          x2 = LayerNorm2(x1 + relu(x1 W1^T + b1) W2^T + b2)
// Reference: geometry_rl/modules/pyg_models/transformer_vanilla.py:30-36,76-92 — nn.TransformerEncoderLayer(d_model = 64,
// nhead = 2, dim_feedforward = 64, dropout = 0, post-LN, ReLU) applied to [S, B, 64]; here tokens are batch-major
// [B][S][64] (the permutes of the reference disappear) and S <= 56 (the shipped config has S = 50).
//
// One CTA per graph at a time (persistent over graphs), 512 threads, everything of a graph in shared memory: the
// reference's ~30 library launches per layer and direction (tiny GEMMs, softmax, LayerNorm, adds) become one.  The
// backward recomputes the layer's forward from its input (the only tensor autograd keeps) and leaves the parameter
// gradients in one partial slot per CTA, accumulated over its graphs in graph order and summed over CTAs in fixed order
// by grl_reduce_partials: deterministic, no atomics.
#include "grl_common.cuh"

namespace grl {

constexpr int kEncD = 64;            // model width
constexpr int kEncHeads = 2;
constexpr int kEncHd = 32;           // head width
constexpr int kEncLdq = 193;         // row stride of the [S][192] q|k|v tile (odd: conflict-free column walks)
constexpr int kEncMaxS = GRL_ENCODER_MAX_TOKENS;
constexpr int kEncThreads = 512;     // 16 warps: 8 row groups x 64 columns in the dense products
constexpr int kEncGroups = kEncThreads / 64;
constexpr int kEncWarps = kEncThreads / 32;

// parameter-gradient slot layout (floats); must match include/grl_b200.h
constexpr int kEgWqkv = 0;                         // [192][64]
constexpr int kEgBqkv = kEgWqkv + 192 * 64;        // [192]
constexpr int kEgWo = kEgBqkv + 192;               // [64][64]
constexpr int kEgBo = kEgWo + 64 * 64;
constexpr int kEgW1 = kEgBo + 64;
constexpr int kEgB1 = kEgW1 + 64 * 64;
constexpr int kEgW2 = kEgB1 + 64;
constexpr int kEgB2 = kEgW2 + 64 * 64;
constexpr int kEgLn1w = kEgB2 + 64;
constexpr int kEgLn1b = kEgLn1w + 64;
constexpr int kEgLn2w = kEgLn1b + 64;
constexpr int kEgLn2b = kEgLn2w + 64;
static_assert(kEgLn2b + 64 == GRL_ENCODER_GRAD_FLOATS, "encoder partial layout");

struct EncBuf {  // float offsets into the dynamic shared memory of a graph with S tokens
  int x, qkv, p, ctx, xh1, x1, h, xh2, g0, dqkv, rstd1, rstd2, total;
  __host__ __device__ static int pad4(int n) { return (n + 3) & ~3; }  // every tile starts 16-byte aligned
  __host__ __device__ EncBuf(int S, bool bwd) {
    int o = 0;
    x = o; o += S * kEncD;
    qkv = o; o += pad4(S * kEncLdq);
    p = o; o += pad4(kEncHeads * S * (S + 1));
    ctx = o; o += S * kEncD;
    xh1 = o; o += S * kEncD;
    x1 = o; o += S * kEncD;      // x1 and xh2 are adjacent: the backward's dP tile [2][S][S+1] overlays them
    xh2 = o; o += S * kEncD;
    h = o; o += S * kEncD;
    rstd1 = o; o += kEncMaxS;
    rstd2 = o; o += kEncMaxS;
    g0 = o; dqkv = o;
    if (bwd) { o += S * kEncD; dqkv = o; o += 3 * S * kEncD; }  // dq | dk | dv, one [S][64] tile each
    total = o;
  }
};

// out[s][n] (+)= sum_k in[s][k] W[n][k] (+ bias[n]);  W row-major [64][ldw] in global memory (L1 / L2 resident).
// Thread (n = tid & 63, g = tid >> 6) keeps row n of W in registers and walks the rows s = g, g + 4, ...
template <bool kAdd>
__device__ __forceinline__ void enc_linear(float* __restrict__ out, int ldo, const float* __restrict__ in, int ldi,
                                           const float* __restrict__ W, int ldw, const float* __restrict__ bias, int S) {
  const int n = threadIdx.x & 63, g = threadIdx.x >> 6;
  float w[kEncD];
#pragma unroll
  for (int k = 0; k < kEncD; k += 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(W + (size_t)n * ldw + k));
    w[k] = v.x; w[k + 1] = v.y; w[k + 2] = v.z; w[k + 3] = v.w;
  }
  const float b = bias ? __ldg(bias + n) : 0.f;
  unsigned long long w2[kEncD / 2];
#pragma unroll
  for (int k = 0; k < kEncD; k += 2) w2[k >> 1] = pack2(w[k], w[k + 1]);
  for (int s = g; s < S; s += kEncGroups) {
    const float4* r = reinterpret_cast<const float4*>(in + s * ldi);  // ldi % 4 == 0 at every call site
    unsigned long long a01 = pack2(0.f, 0.f), a23 = a01;
#pragma unroll
    for (int k4 = 0; k4 < kEncD / 4; ++k4) {
      const float4 v = r[k4];
      a01 = ffma2(pack2(v.x, v.y), w2[2 * k4], a01);
      a23 = ffma2(pack2(v.z, v.w), w2[2 * k4 + 1], a23);
    }
    float a0, a1, a2, a3;
    unpack2(a01, a0, a1);
    unpack2(a23, a2, a3);
    const float v = ((a0 + a2) + (a1 + a3)) + b;
    if (kAdd) out[s * ldo + n] += v; else out[s * ldo + n] = v;
  }
}

// out[s][k] (+)= sum_n g[s][n] W[n][k]   (product with W instead of W^T: thread k keeps COLUMN k of W)
template <bool kAdd>
__device__ __forceinline__ void enc_linear_t(float* __restrict__ out, int ldo, const float* __restrict__ gin, int ldg,
                                             const float* __restrict__ W, int ldw, int S) {
  const int k = threadIdx.x & 63, g = threadIdx.x >> 6;
  float w[kEncD];
#pragma unroll
  for (int n = 0; n < kEncD; ++n) w[n] = __ldg(W + (size_t)n * ldw + k);
  unsigned long long w2[kEncD / 2];
#pragma unroll
  for (int n = 0; n < kEncD; n += 2) w2[n >> 1] = pack2(w[n], w[n + 1]);
  for (int s = g; s < S; s += kEncGroups) {
    const float4* r = reinterpret_cast<const float4*>(gin + s * ldg);  // ldg % 4 == 0 at every call site
    unsigned long long a01 = pack2(0.f, 0.f), a23 = a01;
#pragma unroll
    for (int n4 = 0; n4 < kEncD / 4; ++n4) {
      const float4 v = r[n4];
      a01 = ffma2(pack2(v.x, v.y), w2[2 * n4], a01);
      a23 = ffma2(pack2(v.z, v.w), w2[2 * n4 + 1], a23);
    }
    float a0, a1, a2, a3;
    unpack2(a01, a0, a1);
    unpack2(a23, a2, a3);
    const float v = (a0 + a2) + (a1 + a3);
    if (kAdd) out[s * ldo + k] += v; else out[s * ldo + k] = v;
  }
}

// slot[n][k] += sum_s g[s][n] in[s][k],  slot_b[n] += sum_s g[s][n]   (this CTA's own partial slot in global memory)
__device__ __forceinline__ void enc_wgrad(float* __restrict__ slot, float* __restrict__ slot_b, const float* __restrict__ gin,
                                          int ldg, const float* __restrict__ in, int ldi, int S) {
  constexpr int kPer = kEncD / kEncGroups;                 // 8 columns of `in` per thread
  const int n = threadIdx.x & 63, kg = threadIdx.x >> 6;  // k = kPer kg .. kPer kg + kPer - 1
  unsigned long long acc[kPer / 2];
#pragma unroll
  for (int i = 0; i < kPer / 2; ++i) acc[i] = pack2(0.f, 0.f);
  float bsum = 0.f;
  for (int s = 0; s < S; ++s) {
    const float gv = gin[s * ldg + n];
    const unsigned long long gg = pack2(gv, gv);
    const float4* r = reinterpret_cast<const float4*>(in + s * ldi + kPer * kg);  // ldi % 4 == 0 at every call site
#pragma unroll
    for (int i = 0; i < kPer / 4; ++i) {
      const float4 v = r[i];
      acc[2 * i] = ffma2(gg, pack2(v.x, v.y), acc[2 * i]);
      acc[2 * i + 1] = ffma2(gg, pack2(v.z, v.w), acc[2 * i + 1]);
    }
    bsum += gv;
  }
  float* dst = slot + (size_t)n * kEncD + kPer * kg;
#pragma unroll
  for (int i = 0; i < kPer / 2; ++i) {
    float lo, hi;
    unpack2(acc[i], lo, hi);
    dst[2 * i] += lo;
    dst[2 * i + 1] += hi;
  }
  if (slot_b && kg == 0) slot_b[n] += bsum;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// y = LayerNorm(a + b) over the 64 channels of a row (biased variance, eps 1e-5): xhat and 1/std kept for the backward.
__device__ __forceinline__ void enc_add_layernorm(float* __restrict__ y, float* __restrict__ xhat, float* __restrict__ rstd,
                                                  const float* __restrict__ a, const float* __restrict__ b,
                                                  const float* __restrict__ w, const float* __restrict__ bias, int S) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float w0 = __ldg(w + lane), w1 = __ldg(w + lane + 32), b0 = __ldg(bias + lane), b1 = __ldg(bias + lane + 32);
  for (int s = warp; s < S; s += kEncWarps) {
    const float v0 = a[s * kEncD + lane] + b[s * kEncD + lane], v1 = a[s * kEncD + lane + 32] + b[s * kEncD + lane + 32];
    const float mean = warp_sum(v0 + v1) * (1.0f / 64.0f);
    const float d0 = v0 - mean, d1 = v1 - mean;
    const float var = warp_sum(d0 * d0 + d1 * d1) * (1.0f / 64.0f);
    const float r = 1.0f / sqrtf(var + 1e-5f);
    const float h0 = d0 * r, h1 = d1 * r;
    if (xhat) { xhat[s * kEncD + lane] = h0; xhat[s * kEncD + lane + 32] = h1; }
    if (rstd && lane == 0) rstd[s] = r;
    y[s * kEncD + lane] = h0 * w0 + b0;
    y[s * kEncD + lane + 32] = h1 * w1 + b1;
  }
}

// in place: g <- d(loss)/d(pre-LayerNorm input) from g = d(loss)/d(LayerNorm output); slot_w / slot_b accumulate the
// affine gradients (warp-private partial sums over its rows, combined through shared memory `red[16][128]`).
__device__ __forceinline__ void enc_layernorm_bwd(float* __restrict__ g, const float* __restrict__ xhat,
                                                  const float* __restrict__ rstd, const float* __restrict__ w,
                                                  float* __restrict__ slot_w, float* __restrict__ slot_b,
                                                  float* __restrict__ red, int S) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float w0 = __ldg(w + lane), w1 = __ldg(w + lane + 32);
  float gw0 = 0.f, gw1 = 0.f, gb0 = 0.f, gb1 = 0.f;
  for (int s = warp; s < S; s += kEncWarps) {
    const float g0 = g[s * kEncD + lane], g1 = g[s * kEncD + lane + 32];
    const float h0 = xhat[s * kEncD + lane], h1 = xhat[s * kEncD + lane + 32];
    gw0 = fmaf(g0, h0, gw0); gw1 = fmaf(g1, h1, gw1);
    gb0 += g0; gb1 += g1;
    const float a0 = g0 * w0, a1 = g1 * w1;
    const float m1 = warp_sum(a0 + a1) * (1.0f / 64.0f);
    const float m2 = warp_sum(a0 * h0 + a1 * h1) * (1.0f / 64.0f);
    const float r = rstd[s];
    g[s * kEncD + lane] = r * (a0 - m1 - h0 * m2);
    g[s * kEncD + lane + 32] = r * (a1 - m1 - h1 * m2);
  }
  red[warp * 128 + lane] = gw0; red[warp * 128 + 32 + lane] = gw1;
  red[warp * 128 + 64 + lane] = gb0; red[warp * 128 + 96 + lane] = gb1;
  __syncthreads();
  if (threadIdx.x < 128) {
    float t = 0.f;
#pragma unroll
    for (int wq = 0; wq < kEncWarps; ++wq) t += red[wq * 128 + threadIdx.x];
    if (threadIdx.x < 64) slot_w[threadIdx.x] += t; else slot_b[threadIdx.x - 64] += t;
  }
  __syncthreads();
}

// ---- attention-sized products, 4 x 4 outputs per thread (a scalar-per-output loop spends two shared-memory loads per
// FMA; the tile spends one per two).  Rows / columns past S are clamped for the loads and skipped for the stores. ------
// C[h][i][j] = cs * sum_c A[i][ao + 32 h + c] B[j][bo + 32 h + c],  c < 32      (scores, dP)
__device__ __forceinline__ void enc_tile_nt(float* __restrict__ C, int ldc, const float* __restrict__ A, int lda, int ao,
                                            const float* __restrict__ B, int ldb, int bo, float cs, int S) {
  const int nt = (S + 3) >> 2;
  for (int t = threadIdx.x; t < kEncHeads * nt * nt; t += kEncThreads) {
    const int hh = t / (nt * nt), r = t - hh * nt * nt, it = r / nt, jt = r - it * nt;
    const float* ap[4]; const float* bp[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      ap[a] = A + min(4 * it + a, S - 1) * lda + ao + kEncHd * hh;
      bp[a] = B + min(4 * jt + a, S - 1) * ldb + bo + kEncHd * hh;
    }
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 8
    for (int c = 0; c < kEncHd; ++c) {
      float av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) { av[a] = ap[a][c] * cs; bv[a] = bp[a][c]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (4 * it + a < S && 4 * jt + b < S) C[(hh * S + 4 * it + a) * ldc + 4 * jt + b] = acc[a][b];
  }
}
// C[i][co + n] = cs * sum_j A[h(n)][i][j] B[j][bo + n],  n < 64, h(n) = n / 32      (ctx = P V, dq = dS K)
__device__ __forceinline__ void enc_tile_nn(float* __restrict__ C, int ldc, int co, const float* __restrict__ A, int lda,
                                            const float* __restrict__ B, int ldb, int bo, float cs, int S) {
  const int nt = (S + 3) >> 2;
  for (int t = threadIdx.x; t < nt * 16; t += kEncThreads) {
    const int it = t >> 4, n0 = 4 * (t & 15), hh = n0 >> 5;
    const float* ap[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) ap[a] = A + (hh * S + min(4 * it + a, S - 1)) * lda;
    const float* bp = B + bo + n0;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int j = 0; j < S; ++j) {
      float av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) { av[a] = ap[a][j]; bv[a] = bp[j * ldb + a]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
      if (4 * it + a < S)
#pragma unroll
        for (int b = 0; b < 4; ++b) C[(4 * it + a) * ldc + co + n0 + b] = acc[a][b] * cs;
  }
}
// C[j][co + n] = cs * sum_i A[h(n)][i][j] B[i][bo + n]      (dV = P^T dCTX, dk = dS^T Q)
__device__ __forceinline__ void enc_tile_tn(float* __restrict__ C, int ldc, int co, const float* __restrict__ A, int lda,
                                            const float* __restrict__ B, int ldb, int bo, float cs, int S) {
  const int nt = (S + 3) >> 2;
  for (int t = threadIdx.x; t < nt * 16; t += kEncThreads) {
    const int jt = t >> 4, n0 = 4 * (t & 15), hh = n0 >> 5;
    int jc[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) jc[a] = min(4 * jt + a, S - 1);
    const float* ab = A + hh * S * lda;
    const float* bp = B + bo + n0;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int i = 0; i < S; ++i) {
      float av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) { av[a] = ab[i * lda + jc[a]]; bv[a] = bp[i * ldb + a]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
      if (4 * jt + a < S)
#pragma unroll
        for (int b = 0; b < 4; ++b) C[(4 * jt + a) * ldc + co + n0 + b] = acc[a][b] * cs;
  }
}

// forward of one graph into the shared-memory buffers of `L` (x already staged); y2 (the layer output) -> out_rows.
// With kKeep the tiles the backward needs stay behind (xh1, x1, h, xh2, rstd); without it they are scratch.
__device__ __forceinline__ void enc_forward_graph(float* __restrict__ sm, const EncBuf& L, const GrlEncoderDesc& d, int S,
                                                  float* __restrict__ out_rows) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* X = sm + L.x; float* QKV = sm + L.qkv; float* P = sm + L.p; float* CTX = sm + L.ctx;
  // q | k | v
  enc_linear<false>(QKV, kEncLdq, X, kEncD, d.in_proj_weight, kEncD, d.in_proj_bias, S);
  enc_linear<false>(QKV + 64, kEncLdq, X, kEncD, d.in_proj_weight + 64 * kEncD, kEncD, d.in_proj_bias + 64, S);
  enc_linear<false>(QKV + 128, kEncLdq, X, kEncD, d.in_proj_weight + 128 * kEncD, kEncD, d.in_proj_bias + 128, S);
  __syncthreads();
  // scores: torch scales q by 1/sqrt(head_dim) before the product (F.multi_head_attention_forward)
  const float scale = 0.17677669529663687f;  // 1 / sqrt(32)
  const int ldp = S + 1;
  enc_tile_nt(P, ldp, QKV, kEncLdq, 0, QKV, kEncLdq, 64, scale, S);  // (q * scale) k^T per head
  __syncthreads();
  for (int r = warp; r < kEncHeads * S; r += kEncWarps) {  // softmax over j, one warp per (head, query) row
    float* row = P + r * ldp;
    const float v0 = lane < S ? row[lane] : -INFINITY, v1 = lane + 32 < S ? row[lane + 32] : -INFINITY;
    const float m = warp_max(fmaxf(v0, v1));
    const float e0 = lane < S ? expf(v0 - m) : 0.f, e1 = lane + 32 < S ? expf(v1 - m) : 0.f;
    const float inv = 1.0f / warp_sum(e0 + e1);
    if (lane < S) row[lane] = e0 * inv;
    if (lane + 32 < S) row[lane + 32] = e1 * inv;
  }
  __syncthreads();
  enc_tile_nn(CTX, kEncD, 0, P, ldp, QKV, kEncLdq, 128, 1.0f, S);  // ctx = P v
  __syncthreads();
  float* A = sm + L.h;  // attention output, then the hidden layer
  enc_linear<false>(A, kEncD, CTX, kEncD, d.out_proj_weight, kEncD, d.out_proj_bias, S);
  __syncthreads();
  enc_add_layernorm(sm + L.x1, sm + L.xh1, sm + L.rstd1, X, A, d.norm1_weight, d.norm1_bias, S);
  __syncthreads();
  {  // h = relu(x1 W1^T + b1)
    enc_linear<false>(A, kEncD, sm + L.x1, kEncD, d.linear1_weight, kEncD, d.linear1_bias, S);
    __syncthreads();
    for (int i = tid; i < S * kEncD; i += kEncThreads) A[i] = fmaxf(A[i], 0.f);
    __syncthreads();
  }
  float* F2 = sm + L.ctx;  // ctx is dead in the forward-only pass; the backward recomputes into its own scratch below
  if (out_rows == nullptr) F2 = sm + L.g0;  // backward: keep ctx (needed by dWo), use the gradient scratch
  enc_linear<false>(F2, kEncD, A, kEncD, d.linear2_weight, kEncD, d.linear2_bias, S);
  __syncthreads();
  if (out_rows) {
    enc_add_layernorm(sm + L.xh2, nullptr, nullptr, sm + L.x1, F2, d.norm2_weight, d.norm2_bias, S);
    __syncthreads();
    for (int i = tid; i < S * kEncD / 4; i += kEncThreads)
      reinterpret_cast<float4*>(out_rows)[i] = reinterpret_cast<const float4*>(sm + L.xh2)[i];
  } else {
    float* scratch = sm + L.dqkv;  // LayerNorm output itself is not needed by the backward, only xhat2 and rstd2
    enc_add_layernorm(scratch, sm + L.xh2, sm + L.rstd2, sm + L.x1, F2, d.norm2_weight, d.norm2_bias, S);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kEncThreads, 1) encoder_layer_fwd_kernel(const GrlEncoderDesc d) {
  extern __shared__ __align__(16) float enc_sm[];
  const int S = d.n_tokens;
  const EncBuf L(S, false);
  for (int gidx = blockIdx.x; gidx < d.n_graphs; gidx += gridDim.x) {
    const float4* src = reinterpret_cast<const float4*>(d.x + (size_t)gidx * S * kEncD);
    for (int i = threadIdx.x; i < S * kEncD / 4; i += kEncThreads) reinterpret_cast<float4*>(enc_sm + L.x)[i] = __ldg(src + i);
    __syncthreads();
    enc_forward_graph(enc_sm, L, d, S, d.out + (size_t)gidx * S * kEncD);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kEncThreads, 1) encoder_layer_bwd_kernel(const GrlEncoderDesc d) {
  extern __shared__ __align__(16) float enc_sm[];
  __shared__ float red[kEncWarps * 128];
  const int S = d.n_tokens, tid = threadIdx.x;
  const EncBuf L(S, true);
  float* slot = d.grad_partials + (size_t)blockIdx.x * GRL_ENCODER_GRAD_FLOATS;
  for (int i = tid; i < GRL_ENCODER_GRAD_FLOATS; i += kEncThreads) slot[i] = 0.f;
  __syncthreads();
  float* sm = enc_sm;
  const int ldp = S + 1;
  for (int gidx = blockIdx.x; gidx < d.n_graphs; gidx += gridDim.x) {
    const float4* src = reinterpret_cast<const float4*>(d.x + (size_t)gidx * S * kEncD);
    for (int i = tid; i < S * kEncD / 4; i += kEncThreads) reinterpret_cast<float4*>(sm + L.x)[i] = __ldg(src + i);
    __syncthreads();
    enc_forward_graph(sm, L, d, S, nullptr);  // recompute: qkv, P, ctx, xh1, rstd1, x1, h, xh2, rstd2 are in place
    float* G0 = sm + L.g0;
    const float4* gsrc = reinterpret_cast<const float4*>(d.grad_out + (size_t)gidx * S * kEncD);
    for (int i = tid; i < S * kEncD / 4; i += kEncThreads) reinterpret_cast<float4*>(G0)[i] = __ldg(gsrc + i);
    __syncthreads();
    // LayerNorm2: G0 <- dY2 (gradient of x1 + ff)
    enc_layernorm_bwd(G0, sm + L.xh2, sm + L.rstd2, d.norm2_weight, slot + kEgLn2w, slot + kEgLn2b, red, S);
    // feed-forward: T0 = dY2 W2 (into xh2, dead), gated by relu;  dW2 += dY2^T h
    float* T0 = sm + L.xh2;
    float* H = sm + L.h;
    enc_linear_t<false>(T0, kEncD, G0, kEncD, d.linear2_weight, kEncD, S);
    enc_wgrad(slot + kEgW2, slot + kEgB2, G0, kEncD, H, kEncD, S);
    __syncthreads();
    for (int i = tid; i < S * kEncD; i += kEncThreads) T0[i] = H[i] > 0.f ? T0[i] : 0.f;
    __syncthreads();
    // dX1 = dY2 + T0 W1 (accumulated into G0);  dW1 += T0^T x1
    enc_linear_t<true>(G0, kEncD, T0, kEncD, d.linear1_weight, kEncD, S);
    enc_wgrad(slot + kEgW1, slot + kEgB1, T0, kEncD, sm + L.x1, kEncD, S);
    __syncthreads();
    // LayerNorm1: G0 <- dY1 (gradient of x + attention output)
    enc_layernorm_bwd(G0, sm + L.xh1, sm + L.rstd1, d.norm1_weight, slot + kEgLn1w, slot + kEgLn1b, red, S);
    // out projection: dCTX = dY1 Wo (into xh1, dead);  dWo += dY1^T ctx
    float* DC = sm + L.xh1;
    enc_linear_t<false>(DC, kEncD, G0, kEncD, d.out_proj_weight, kEncD, S);
    enc_wgrad(slot + kEgWo, slot + kEgBo, G0, kEncD, sm + L.ctx, kEncD, S);
    __syncthreads();
    // attention: dP[h][i][j] = sum_c dCTX[i][c] v[j][c]  (overlays x1 | xh2, both dead)
    float* QKV = sm + L.qkv; float* P = sm + L.p; float* DP = sm + L.x1; float* DQ = sm + L.dqkv;
    enc_tile_nt(DP, ldp, DC, kEncD, 0, QKV, kEncLdq, 128, 1.0f, S);            // dP = dCTX v^T per head
    enc_tile_tn(DQ, kEncD, 2 * S * kEncD, P, ldp, DC, kEncD, 0, 1.0f, S);     // dV = P^T dCTX
    __syncthreads();
    {  // softmax backward in place: dS = P (dP - sum_j P dP), one warp per (head, query) row
      const int warp = tid >> 5, lane = tid & 31;
      for (int r = warp; r < kEncHeads * S; r += kEncWarps) {
        float* dp = DP + r * ldp; const float* p = P + r * ldp;
        const float p0 = lane < S ? p[lane] : 0.f, p1 = lane + 32 < S ? p[lane + 32] : 0.f;
        const float d0 = lane < S ? dp[lane] : 0.f, d1 = lane + 32 < S ? dp[lane + 32] : 0.f;
        const float dot = warp_sum(p0 * d0 + p1 * d1);
        if (lane < S) dp[lane] = p0 * (d0 - dot);
        if (lane + 32 < S) dp[lane + 32] = p1 * (d1 - dot);
      }
    }
    __syncthreads();
    {  // dq = scale dS k,  dk = scale dS^T q
      const float scale = 0.17677669529663687f;
      enc_tile_nn(DQ, kEncD, 0, DP, ldp, QKV, kEncLdq, 64, scale, S);
      enc_tile_tn(DQ, kEncD, S * kEncD, DP, ldp, QKV, kEncLdq, 0, scale, S);
    }
    __syncthreads();
    // in projection: dX = dY1 + dQ Wq + dK Wk + dV Wv;  dWqkv += dQKV^T x
    float* X = sm + L.x;
    for (int part = 0; part < 3; ++part) {  // q, k, v
      const float* dpart = DQ + part * S * kEncD;
      enc_linear_t<true>(G0, kEncD, dpart, kEncD, d.in_proj_weight + part * 64 * kEncD, kEncD, S);
      enc_wgrad(slot + kEgWqkv + part * 64 * kEncD, slot + kEgBqkv + part * 64, dpart, kEncD, X, kEncD, S);
      __syncthreads();
    }
    float4* gdst = reinterpret_cast<float4*>(d.grad_x + (size_t)gidx * S * kEncD);
    for (int i = tid; i < S * kEncD / 4; i += kEncThreads) gdst[i] = reinterpret_cast<const float4*>(G0)[i];
    __syncthreads();
  }
}

}  // namespace grl

extern "C" {

static int enc_check(const GrlEncoderDesc* d, const char* who) {
  GRL_REQUIRE(d, GRL_EINVAL, "%s: null descriptor", who);
  GRL_REQUIRE(d->n_graphs > 0 && d->n_tokens > 0, GRL_EINVAL, "%s: n_graphs=%d n_tokens=%d", who, d->n_graphs, d->n_tokens);
  GRL_REQUIRE(d->n_tokens <= GRL_ENCODER_MAX_TOKENS, GRL_EUNSUPPORTED, "%s: %d tokens per graph, at most %d are supported", who,
              d->n_tokens, GRL_ENCODER_MAX_TOKENS);
  GRL_REQUIRE(d->x && d->in_proj_weight && d->in_proj_bias && d->out_proj_weight && d->out_proj_bias && d->linear1_weight &&
                  d->linear1_bias && d->linear2_weight && d->linear2_bias && d->norm1_weight && d->norm1_bias &&
                  d->norm2_weight && d->norm2_bias, GRL_EINVAL, "%s: null pointer", who);
  return GRL_OK;
}

int grl_encoder_layer_fwd(const GrlEncoderDesc* d, grl_stream_t stream) {
  const int rc = enc_check(d, "grl_encoder_layer_fwd");
  if (rc != GRL_OK) return rc;
  GRL_REQUIRE(d->out, GRL_EINVAL, "grl_encoder_layer_fwd: out is null");
  const int smem = grl::EncBuf(d->n_tokens, false).total * 4;
  // the attribute is set once per (kernel, device): ask for the largest token count the kernel supports
  if (grl::ensure_dynamic_smem((const void*)grl::encoder_layer_fwd_kernel, grl::EncBuf(GRL_ENCODER_MAX_TOKENS, false).total * 4) != GRL_OK)
    return GRL_ECUDA;
  int grid = grl::sm_count();
  if (grid > d->n_graphs) grid = d->n_graphs;
  grl::encoder_layer_fwd_kernel<<<grid, grl::kEncThreads, smem, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_encoder_layer_fwd");
}

int grl_encoder_layer_bwd(const GrlEncoderDesc* d, grl_stream_t stream) {
  const int rc = enc_check(d, "grl_encoder_layer_bwd");
  if (rc != GRL_OK) return rc;
  GRL_REQUIRE(d->grad_out && d->grad_x && d->grad_partials && d->n_partials > 0 && d->n_partials <= d->n_graphs, GRL_EINVAL,
              "grl_encoder_layer_bwd: null pointer or n_partials=%d outside [1, n_graphs]", d->n_partials);
  const int smem = grl::EncBuf(d->n_tokens, true).total * 4;
  if (grl::ensure_dynamic_smem((const void*)grl::encoder_layer_bwd_kernel, grl::EncBuf(GRL_ENCODER_MAX_TOKENS, true).total * 4) != GRL_OK)
    return GRL_ECUDA;
  grl::encoder_layer_bwd_kernel<<<d->n_partials, grl::kEncThreads, smem, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_encoder_layer_bwd");
}

}  // extern "C"
