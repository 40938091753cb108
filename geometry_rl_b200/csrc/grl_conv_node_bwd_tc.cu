// Backward of the node half of the fibre-bundle convolution, tensor-core path.  Two kernels:
//
//  (1) fbconv_node_bwd_tc2_kernel — reads the pre-LayerNorm tensor x2 saved by the forward kernel, recomputes
//      LayerNorm + GEMM1, then on the tensor cores
//        pre  = y W1^T + b1                 (recompute)      h = GELU(pre), dG = GELU'(pre)
//        gH   = gZ W2                       gPre = gH * dG
//        gY   = gPre W1                     dW1 += gPre^T y       dW2^T += h^T gZ
//      and on the CUDA cores the LayerNorm backward (gY -> g_x2, written to HBM), gb2 / g_ln_g / g_ln_b.
//      The two weight gradients (and gb1, as an extra column of dW1) live in TMEM for the whole kernel, accumulated
//      over every tile of the CTA by the MMA itself, and are written once to this CTA's partial slot:
//      deterministic, no atomics.
//  (2) fbconv_fiber_bwd_kernel    — fp32: g_x1[o] = 1/16 sum_p g_x2[p] fk[o][p],  g_fk += 1/16 x1[o] g_x2[p],
//      g_bias += g_x2.
//
// Reference: autograd of ponita/conv.py:88-114 (FiberBundleConv.forward node part).
// Operand images are [chunk][row][8] 16-bit (grl_tc.cuh); the SAME image is read K-major (activation x weight)
// and MN-major (weight gradients X^T Y, products with W instead of W^T) so nothing is ever transposed.
#include "grl_common.cuh"
#include "grl_tc.cuh"

namespace grl {

// node partial slot layout (must match grl_conv_node.cu / include/grl_b200.h)
constexpr int kPGW1 = 0;
constexpr int kPGB1 = kPGW1 + kH * kC;
constexpr int kPGW2 = kPGB1 + kH;
constexpr int kPGB2 = kPGW2 + kC * kH;
constexpr int kPGLNG = kPGB2 + kC;
constexpr int kPGLNB = kPGLNG + kC;
constexpr int kPGBIAS = kPGLNB + kC;
constexpr int kPGFK = kPGBIAS + kC;
static_assert(kPGFK + kO * kO * kC == GRL_NODE_GRAD_FLOATS, "partial layout");

// ---------------------------------------------------------------------------------------------------
// Kernel (1): 512 threads (16 warps) on one 128-row tile, fp16 operands.
//   * Every operand is fp16 (11-bit mantissa instead of bf16's 8): activations y, h and the weights are O(1); the
//     gradients (grad_out, gPre) are multiplied by a power of two derived from grl_absmax(grad_out) so that
//     max |g| lies in [32, 64) (grl_tc.cuh grad_scale_from_amax) and the factor is removed exactly in the epilogues.
//     tcgen05 kind::f16 needs A and B in ONE format (fp16 x bf16 is an illegal instruction on B200), which is why
//     the gradients cannot simply stay bf16 next to fp16 weights.
//   * b1 rides in the contraction (K = 80: y[:, 64:66] = 1, W1[:, 64:66] = fp16 (hi, lo) split of b1) and the same
//     ones-columns turn the dW1 MMA (N = 80) into the b1 gradient: no bias adds, no 64-value column butterfly.
//   * GELU and GELU' are evaluated two elements per instruction in packed fp16 (gelu_h2); h and dG stay packed.
//   * the x2 and grad_out tiles of the NEXT tile are fetched with cp.async as soon as this tile's last MMA has
//     retired, and pulled into L2 one tile earlier still.
//   * no fibre recompute: the forward kernel saves the pre-LayerNorm tensor x2 (GrlConvDesc.x2) and this kernel
//     streams it back with cp.async (the fibre phase was 20 % of the stall samples of the recomputing version);
//   * half 0 never waits on its gY / dW MMAs: they are issued behind pre(1) and retire under the next epilogue (the
//     tensor pipe completes MMAs in issue order, so a later wait covers them).
// TMEM columns: D 0..127 | gY 128..191 | dW1 192..351 (2 x 80) | dW2^T 352..479 (2 x 64).
// ---------------------------------------------------------------------------------------------------
constexpr int kNB2Threads = 512;
constexpr int kKb = 80;     // GEMM1 contraction / dW1 width: 64 channels + 2 ones columns + 14 zero columns
constexpr int kLDX2 = 68;   // X2 row stride (floats): conflict-free for the row-per-lane LayerNorm reads

struct NodeBwd2Smem {
  __half W1h[kH * kKb];     // [10 chunks][256 rows k'][8]; chunk 8 = (b1 hi, b1 lo, 0 ...), chunk 9 = 0
  __half W2h[kC * kH];      // [32 chunks k'][64 rows n][8]
  __half A1[kTM * kKb];     // y: [10 chunks][128 rows][8]; chunk 8 = (1, 1, 0 ...), chunk 9 = 0 (written once)
  __half GZh[kTM * kC];     // scaled grad_out [8 chunks][128 rows][8]
  union {
    struct {
      float X2[kTM * kLDX2];    // pre-LayerNorm x2 tile (saved by the forward kernel), phase B only
      float GZf[kTM * kLDX2];   // grad_out tile (fp32), phase B only
    } in;
    struct {
      __half A2h[kTM * 128];    // h           half: [16 chunks][128 rows][8]
      __half AP[kTM * 128];     // scaled gPre half: [16 chunks][128 rows][8]
    } h;
  } u;
  float bias[kC], lng[kC], lnb[kC];
  float rs[2][4][kTM];          // row partial sums exchanged between the four column groups
  float acc_gb2[4][kC], acc_glng[4][kC], acc_glnb[4][kC];  // per lane-quarter column sums
  uint64_t bar[3];
  uint32_t tmem_base;
};

constexpr uint32_t kCol2D = 0, kCol2GY = 128, kCol2DW1 = 192, kCol2DW2 = 352;

// column sums over the 32 lanes of 16 per-lane values: on return every lane l holds the total of index l >> 1 in v[0]
__device__ __forceinline__ void warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    const int half = 8 >> step;
    const int bit = 16 >> step;
    const bool upper = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = upper ? v[i + half] : v[i];
      const float send = upper ? v[i] : v[i + half];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

__global__ void __launch_bounds__(kNB2Threads, 1) fbconv_node_bwd_tc2_kernel(const GrlConvDesc d) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  NodeBwd2Smem& s = *reinterpret_cast<NodeBwd2Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, cg = warp >> 2;   // TMEM lane quarter / column group (0..3) of this warp
  const int row = 32 * q + lane;            // tile row owned in the LayerNorm and epilogue phases

  if (tid == 0) {
    tc::mbar_init(&s.bar[0], 1);
    tc::mbar_init(&s.bar[1], 1);
    tc::mbar_init(&s.bar[2], 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&s.tmem_base, 512);
  tc::stage_weight_f16(s.W1h, d.w1, kH, kC, kC);   // chunks 0..7
  tc::stage_weight_f16(s.W2h, d.w2, kC, kH, kH);
  for (int n = tid; n < kH; n += kNB2Threads) {     // chunk 8: fp16 (hi, lo) split of b1; chunk 9: zero
    const float b = d.b1[n];
    const __half hi = __float2half_rn(b);
    const __half lo = __float2half_rn(b - __half2float(hi));
    const __half2 p0 = __halves2half2(hi, lo);
    *reinterpret_cast<uint4*>(s.W1h + ((size_t)8 * kH + n) * 8) = make_uint4(*reinterpret_cast<const uint32_t*>(&p0), 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(s.W1h + ((size_t)9 * kH + n) * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid < kTM) {  // ones columns of y (fp16 1.0 = 0x3C00), persistent
    *reinterpret_cast<uint4*>(s.A1 + ((size_t)8 * kTM + tid) * 8) = make_uint4(0x3C003C00u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(s.A1 + ((size_t)9 * kTM + tid) * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid < kC) { s.bias[tid] = d.bias[tid]; s.lng[tid] = d.ln_g[tid]; s.lnb[tid] = d.ln_b[tid]; }
  for (int i = tid; i < 4 * kC; i += kNB2Threads) {
    (&s.acc_gb2[0][0])[i] = 0.f; (&s.acc_glng[0][0])[i] = 0.f; (&s.acc_glnb[0][0])[i] = 0.f;
  }
  const float gscale = d.grad_amax ? tc::grad_scale_from_amax(__ldg(d.grad_amax)) : 1.0f;
  const float inv_gscale = 1.0f / gscale;  // exact: gscale is a power of two
  const int n_tiles = (d.n_dst + kTE - 1) / kTE;
  int tile = blockIdx.x;
  auto stage_x2 = [&](int t) {  // x2 and grad_out rows of tile t -> X2 / GZf (row stride kLDX2), 16-byte cp.async pieces
    const int cnt = min(kTE, d.n_dst - t * kTE);
    const float* src = d.x2 + (size_t)t * kTE * kRow;
    const float* gsrc = d.grad_out + (size_t)t * kTE * kRow;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int f = tid + kNB2Threads * i;  // float4 index 0..2047 of the [128][64] tile
      const int so = (f >> 4) * kLDX2 + 4 * (f & 15);
      if ((f >> 8) < cnt) {
        cp_async16(s.u.in.X2 + so, src + 4 * f);
        cp_async16(s.u.in.GZf + so, gsrc + 4 * f);
      } else {
        *reinterpret_cast<float4*>(s.u.in.X2 + so) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(s.u.in.GZf + so) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  if (tile < n_tiles) {
    stage_x2(tile);
    cp_async_commit();
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = s.tmem_base;
  const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
  const uint32_t a1 = tc::smem_u32(s.A1), gz = tc::smem_u32(s.GZh), a2 = tc::smem_u32(s.u.h.A2h), ap = tc::smem_u32(s.u.h.AP);
  const uint32_t w1 = tc::smem_u32(s.W1h), w2 = tc::smem_u32(s.W2h);
  uint32_t par0 = 0, par1 = 0, par2 = 0;
  bool first_tile = true;

  for (; tile < n_tiles; tile += gridDim.x) {
    const int n0 = tile * kTE;
    const int node = n0 + (row >> 4);
    const bool live = node < d.n_dst;
    const size_t roff = (size_t)(live ? node : 0) * kRow + (row & 15) * kC + 16 * cg;  // this thread's 16 channels
    if (tid == 0) {  // next tile's x1 / grad_out -> L2
      const int nt = tile + gridDim.x;
      if (nt < n_tiles) {
        const uint32_t bytes = (uint32_t)min(kTE, d.n_dst - nt * kTE) * kRow * 4u;
        tc::prefetch_l2(d.x2 + (size_t)nt * kTE * kRow, bytes);
        tc::prefetch_l2(d.grad_out + (size_t)nt * kTE * kRow, bytes);
      }
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- B: LayerNorm forward (thread = row x 16-channel group cg) + scaled grad_out -> fp16 operand -------
    float xh[16];  // x-hat of this thread's 16 channels, kept for the LayerNorm backward
    float rstd;
    {
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = ld4(s.u.in.X2 + row * kLDX2 + 16 * cg + 4 * i);
        xh[4 * i] = v.x; xh[4 * i + 1] = v.y; xh[4 * i + 2] = v.z; xh[4 * i + 3] = v.w;
        sum += (v.x + v.y) + (v.z + v.w);
      }
      s.rs[0][cg][row] = sum;
      __syncthreads();
      const float mean = ((s.rs[0][0][row] + s.rs[0][1][row]) + (s.rs[0][2][row] + s.rs[0][3][row])) * (1.0f / 64.0f);
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { xh[i] -= mean; sq = fmaf(xh[i], xh[i], sq); }
      s.rs[1][cg][row] = sq;
      __syncthreads();
      rstd = rsqrtf(((s.rs[1][0][row] + s.rs[1][1][row]) + (s.rs[1][2][row] + s.rs[1][3][row])) * (1.0f / 64.0f) + 1e-5f);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = 16 * cg + 8 * i + e;
          xh[8 * i + e] *= rstd;
          y[e] = xh[8 * i + e] * s.lng[c] + s.lnb[c];
        }
        *reinterpret_cast<uint4*>(s.A1 + ((size_t)(2 * cg + i) * kTM + row) * 8) = tc::pack8_h(y);
      }
      // grad_out row piece: column sums (gb2) of the raw values, then scaled -> GZh
      float g16[16], gs[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 gv = ld4(s.u.in.GZf + row * kLDX2 + 16 * cg + 4 * i);
        g16[4 * i] = gv.x; g16[4 * i + 1] = gv.y; g16[4 * i + 2] = gv.z; g16[4 * i + 3] = gv.w;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) gs[i] = g16[i] * gscale;
      *reinterpret_cast<uint4*>(s.GZh + ((size_t)(2 * cg) * kTM + row) * 8) = tc::pack8_h(gs);
      *reinterpret_cast<uint4*>(s.GZh + ((size_t)(2 * cg + 1) * kTM + row) * 8) = tc::pack8_h(gs + 8);
      warp_colsum16(g16, lane);
      if ((lane & 1) == 0) s.acc_gb2[q][16 * cg + (lane >> 1)] += g16[0];
    }

    // ---- C: the two hidden halves -------------------------------------------------------------------------
#pragma unroll 1
    for (int h2 = 0; h2 < 2; ++h2) {
      tc::fence_async_smem();
      tc::tc_fence_before();
      __syncthreads();
      if (tid == 0) {  // pre = [y | 1 1 0..] [W1 | b1]^T (rows 128 h2 .. of W1)
        tc::tc_fence_after();
        tc::issue_mma(tmem + kCol2D, tc::view_k(a1, kTM), tc::view_k(w1 + 128 * h2 * 16, kH),
                      tc::idesc_f16_ex(128, 128, 0, 0, 0, 0), kKb / 16, false);
        tc::mma_commit(&s.bar[0]);
        if (h2 == 1) {
          // half 0's gY / dW1 are issued BEHIND pre(1): the GELU epilogue below only waits for pre(1), and these two
          // (576 tensor-pipe cycles) run underneath it.  They read AP(0) / A1, which nobody writes before the wait on
          // gH(1) (bar[1]) that, by in-order completion, also covers them.
          tc::issue_mma(tmem + kCol2GY, tc::view_k(ap, kTM), tc::view_mn(w1, kH), tc::idesc_f16_ex(128, 64, 0, 1, 0, 0),
                        128 / 16, false);
          tc::issue_mma(tmem + kCol2DW1, tc::view_mn(ap, kTM), tc::view_mn(a1, kTM), tc::idesc_f16_ex(128, kKb, 1, 1, 0, 0),
                        kTM / 16, !first_tile);
        }
      }
      tc::mbar_wait(&s.bar[0], par0);
      par0 ^= 1u;
      tc::tc_fence_after();
      __half2 dG[16];  // GELU'(pre) of this thread's 32 columns of the half, packed
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c0 = 32 * cg + 16 * i;  // column inside the half
        float v[16];
        tc::tmem_ld16(lane_addr + kCol2D + c0, v);
        __half2 h0[4], h1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          tc::gelu_h2(__floats2half2_rn(v[2 * e], v[2 * e + 1]), h0[e], dG[8 * i + e]);
          tc::gelu_h2(__floats2half2_rn(v[8 + 2 * e], v[8 + 2 * e + 1]), h1[e], dG[8 * i + 4 + e]);
        }
        *reinterpret_cast<uint4*>(s.u.h.A2h + ((size_t)(c0 >> 3) * kTM + row) * 8) = tc::pack_h8(h0);
        *reinterpret_cast<uint4*>(s.u.h.A2h + ((size_t)((c0 >> 3) + 1) * kTM + row) * 8) = tc::pack_h8(h1);
      }
      tc::fence_async_smem();
      tc::tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc::tc_fence_after();
        // gH' = gZ' W2[:, half]   (W2h read MN-major: N = k', K = n)
        tc::issue_mma(tmem + kCol2D, tc::view_k(gz, kTM), tc::view_mn(w2 + (16 * h2) * (kC * 16), kC),
                      tc::idesc_f16_ex(128, 128, 0, 1, 0, 0), kC / 16, false);
        tc::mma_commit(&s.bar[1]);  // the epilogue below only needs gH': dW2 runs underneath it
        // dW2'^T[k'][n] += h^T gZ'   (both MN-major, K = tile rows)
        tc::issue_mma(tmem + kCol2DW2 + 64 * h2, tc::view_mn(a2, kTM), tc::view_mn(gz, kTM),
                      tc::idesc_f16_ex(128, 64, 1, 1, 0, 0), kTM / 16, !first_tile);
      }
      tc::mbar_wait(&s.bar[1], par1);
      par1 ^= 1u;
      tc::tc_fence_after();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c0 = 32 * cg + 16 * i;
        float v[16];
        tc::tmem_ld16(lane_addr + kCol2D + c0, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 da = __half22float2(dG[8 * i + e]), db = __half22float2(dG[8 * i + 4 + e]);
          v[2 * e] *= da.x;
          v[2 * e + 1] *= da.y;
          v[8 + 2 * e] *= db.x;
          v[8 + 2 * e + 1] *= db.y;
        }
        *reinterpret_cast<uint4*>(s.u.h.AP + ((size_t)(c0 >> 3) * kTM + row) * 8) = tc::pack8_h(v);
        *reinterpret_cast<uint4*>(s.u.h.AP + ((size_t)((c0 >> 3) + 1) * kTM + row) * 8) = tc::pack8_h(v + 8);
      }
      if (h2 == 0) continue;  // half 0's gY / dW1 are issued at the top of the next iteration, behind pre(1)
      tc::fence_async_smem();
      tc::tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc::tc_fence_after();
        // gY' += gPre' W1[half]   (W1h read MN-major: N = c, K = k')
        tc::issue_mma(tmem + kCol2GY, tc::view_k(ap, kTM), tc::view_mn(w1 + 128 * 16, kH),
                      tc::idesc_f16_ex(128, 64, 0, 1, 0, 0), 128 / 16, true);
        // [dW1' | gb1'][k'][c] += gPre'^T [y | 1 1 0..]
        tc::issue_mma(tmem + kCol2DW1 + kKb, tc::view_mn(ap, kTM), tc::view_mn(a1, kTM),
                      tc::idesc_f16_ex(128, kKb, 1, 1, 0, 0), kTM / 16, !first_tile);
        tc::mma_commit(&s.bar[2]);
      }
    }
    tc::mbar_wait(&s.bar[2], par2);  // everything of this tile is complete: gY is final, the operand buffers are free
    par2 ^= 1u;
    tc::tc_fence_after();

    // A2h / AP (= X2) are free: prefetch the next tile's x2
    {
      const int nt = tile + gridDim.x;
      if (nt < n_tiles) {
        stage_x2(nt);
        cp_async_commit();
      }
    }

    // ---- D: LayerNorm backward: gY (TMEM) -> g_x2 (HBM), g_ln_g, g_ln_b --------------------------------------
    {
      float gy[32];  // [0..15] = gY, [16..31] = gY * x-hat (second half filled below for the column sums)
      {
        float v[16];
        tc::tmem_ld16(lane_addr + kCol2GY + 16 * cg, v);
#pragma unroll
        for (int e = 0; e < 16; ++e) gy[e] = v[e] * inv_gscale;
      }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float hx = gy[i] * s.lng[16 * cg + i];
        s1 += hx;
        s2 = fmaf(hx, xh[i], s2);
      }
      s.rs[0][cg][row] = s1;
      s.rs[1][cg][row] = s2;
      __syncthreads();
      const float m1 = ((s.rs[0][0][row] + s.rs[0][1][row]) + (s.rs[0][2][row] + s.rs[0][3][row])) * (1.0f / 64.0f);
      const float m2 = ((s.rs[1][0][row] + s.rs[1][1][row]) + (s.rs[1][2][row] + s.rs[1][3][row])) * (1.0f / 64.0f);
      float* gdst = d.grad_x2 + roff;  // consumed by kernel (2)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = 4 * i + e;
          const float hx = gy[c] * s.lng[16 * cg + c];
          o[e] = rstd * (hx - m1 - xh[c] * m2);
        }
        if (live) st4(gdst + 4 * i, make_float4(o[0], o[1], o[2], o[3]));
      }
      // column sums over this warp's 32 rows: g_ln_b = sum gy (lanes 0..15), g_ln_g = sum gy * xhat (lanes 16..31)
#pragma unroll
      for (int i = 0; i < 16; ++i) gy[16 + i] = gy[i] * xh[i];
      tc::warp_colsum<32>(gy, lane);
      if (lane < 16) s.acc_glnb[q][16 * cg + lane] += gy[0];
      else s.acc_glng[q][16 * cg + lane - 16] += gy[0];
    }
    tc::tc_fence_before();
    first_tile = false;
  }

  // ---- write this CTA's partial slot (weight gradients unscaled here) ------------------------------------------
  cp_async_wait_all();
  __syncthreads();
  tc::tc_fence_after();
  float* P = d.node_grad_partials + (size_t)blockIdx.x * GRL_NODE_GRAD_FLOATS;
#pragma unroll 1
  for (int h2 = 0; h2 < 2; ++h2) {
    const int c0 = 16 * cg;  // warp (q, cg): rows 128 h2 + row, columns c0 .. c0 + 15
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = 0.f;  // CTA without work: TMEM was never written
    if (!first_tile) tc::tmem_ld16(lane_addr + kCol2DW1 + kKb * h2 + c0, v);
    float* p1 = P + kPGW1 + (size_t)(128 * h2 + row) * kC + c0;
#pragma unroll
    for (int e = 0; e < 16; e += 4)
      st4(p1 + e, make_float4(v[e] * inv_gscale, v[e + 1] * inv_gscale, v[e + 2] * inv_gscale, v[e + 3] * inv_gscale));
    if (cg == 0) {  // gb1 = the first ones column of the dW1 accumulator
      float b[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) b[e] = 0.f;
      if (!first_tile) tc::tmem_ld16(lane_addr + kCol2DW1 + kKb * h2 + 64, b);
      P[kPGB1 + 128 * h2 + row] = b[0] * inv_gscale;
    }
    if (!first_tile) tc::tmem_ld16(lane_addr + kCol2DW2 + 64 * h2 + c0, v);
#pragma unroll
    for (int e = 0; e < 16; ++e) P[kPGW2 + (size_t)(c0 + e) * kH + 128 * h2 + row] = v[e] * inv_gscale;
  }
  if (tid < kC) {
    P[kPGB2 + tid] = ((s.acc_gb2[0][tid] + s.acc_gb2[1][tid]) + s.acc_gb2[2][tid]) + s.acc_gb2[3][tid];
    P[kPGLNG + tid] = ((s.acc_glng[0][tid] + s.acc_glng[1][tid]) + s.acc_glng[2][tid]) + s.acc_glng[3][tid];
    P[kPGLNB + tid] = ((s.acc_glnb[0][tid] + s.acc_glnb[1][tid]) + s.acc_glnb[2][tid]) + s.acc_glnb[3][tid];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// (2) fibre convolution backward, fp32: a pure streaming pass (reads g_x2 and x1, writes g_x1: 12 KB per node).
// 512 threads; thread (channel c, orientation pair op) keeps fk[2 o][16 p] and its g_fk accumulators in registers.
// Tiles of 4 nodes are staged with 16-byte cp.async into a double-buffered shared-memory ring (64 KB in flight per
// SM) so the HBM latency is covered without relying on occupancy.
constexpr int kFiberThreads = 512;
constexpr int kFibTile = 4;
struct FiberBwdSmem {
  float G[2][kFibTile * kRow];
  float X[2][kFibTile * kRow];
};

__device__ __forceinline__ void fiber_stage(FiberBwdSmem& s, int buf, const GrlConvDesc& d, int tile) {
  const int n0 = tile * kFibTile;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int f = threadIdx.x + kFiberThreads * i;  // float4 index over [2 arrays][4 nodes][256 float4]
    const int arr = f >> 10, idx = f & 1023, node = n0 + (idx >> 8);
    float* dst = (arr ? s.X[buf] : s.G[buf]) + 4 * idx;
    if (node < d.n_dst) cp_async16(dst, (arr ? d.x1 : d.grad_x2) + (size_t)n0 * kRow + 4 * idx);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(kFiberThreads, 1) fbconv_fiber_bwd_kernel(const GrlConvDesc d) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FiberBwdSmem& s = *reinterpret_cast<FiberBwdSmem*>(smem_raw);
  const int tid = threadIdx.x, c = tid & 63, op = tid >> 6;
  float fk[2][kO], gfk[2][kO];
#pragma unroll
  for (int oi = 0; oi < 2; ++oi)
#pragma unroll
    for (int p = 0; p < kO; ++p) {
      fk[oi][p] = __ldg(d.fiber_kernel + ((size_t)((2 * op + oi) * kO + p)) * kC + c) * 0.0625f;
      gfk[oi][p] = 0.f;
    }
  float gbias = 0.f;
  const int n_tiles = (d.n_dst + kFibTile - 1) / kFibTile;
  int tile = blockIdx.x, buf = 0;
  if (tile < n_tiles) fiber_stage(s, 0, d, tile);
  cp_async_commit();
  for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
    const int nt = tile + gridDim.x;
    if (nt < n_tiles) fiber_stage(s, buf ^ 1, d, nt);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kFibTile; ++j) {
      const int node = tile * kFibTile + j;
      const float* G = s.G[buf] + j * kRow + c;
      const float* X = s.X[buf] + j * kRow + c;
      float g2[kO];
#pragma unroll
      for (int p = 0; p < kO; ++p) g2[p] = G[p * kC];
#pragma unroll
      for (int oi = 0; oi < 2; ++oi) {
        const float x1v = X[(2 * op + oi) * kC];
        float a = 0.f;
#pragma unroll
        for (int p = 0; p < kO; ++p) {
          a = fmaf(g2[p], fk[oi][p], a);
          gfk[oi][p] = fmaf(x1v, g2[p], gfk[oi][p]);
        }
        if (node < d.n_dst) d.grad_x1[(size_t)node * kRow + (2 * op + oi) * kC + c] = a;
      }
      if (op == 0) {
#pragma unroll
        for (int p = 0; p < kO; ++p) gbias += g2[p];
      }
    }
    __syncthreads();  // everyone done with `buf` before the next iteration's prefetch overwrites it
  }
  float* P = d.node_grad_partials + (size_t)blockIdx.x * GRL_NODE_GRAD_FLOATS;
#pragma unroll
  for (int oi = 0; oi < 2; ++oi)
#pragma unroll
    for (int p = 0; p < kO; ++p) P[kPGFK + ((size_t)((2 * op + oi) * kO + p)) * kC + c] = gfk[oi][p] * 0.0625f;
  if (op == 0) P[kPGBIAS + c] = gbias;
}

}  // namespace grl

extern "C" int grl_fbconv_node_bwd_tc(const GrlConvDesc* d, grl_stream_t stream) {
  GRL_REQUIRE(d && d->n_dst > 0, GRL_EINVAL, "grl_fbconv_node_bwd_tc: bad descriptor");
  GRL_REQUIRE(d->x1 && d->fiber_kernel && d->bias && d->ln_g && d->ln_b && d->w1 && d->b1 && d->w2 && d->grad_out &&
                  d->grad_x1 && d->grad_x2 && d->node_grad_partials && d->n_partials_node > 0, GRL_EINVAL,
              "grl_fbconv_node_bwd_tc: null pointer");
  {
    GRL_REQUIRE(d->x2, GRL_EINVAL, "grl_fbconv_node_bwd_tc: x2 (saved by grl_fbconv_node_fwd_tc) is required");
    const int smem = (int)sizeof(grl::NodeBwd2Smem);
    if (grl::ensure_dynamic_smem((const void*)grl::fbconv_node_bwd_tc2_kernel, smem) != GRL_OK) return GRL_ECUDA;
    grl::fbconv_node_bwd_tc2_kernel<<<d->n_partials_node, grl::kNB2Threads, smem, (cudaStream_t)stream>>>(*d);
  }
  int rc = grl::check_launch("grl_fbconv_node_bwd_tc (mlp)");
  if (rc != GRL_OK) return rc;
  const int smem2 = (int)sizeof(grl::FiberBwdSmem);
  if (grl::ensure_dynamic_smem((const void*)grl::fbconv_fiber_bwd_kernel, smem2) != GRL_OK) return GRL_ECUDA;
  grl::fbconv_fiber_bwd_kernel<<<d->n_partials_node, grl::kFiberThreads, smem2, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_fbconv_node_bwd_tc (fibre)");
}
