// Backward of the node half of the fibre-bundle convolution, tensor-core path.  Two kernels:
//
//  (1) fbconv_node_bwd_tc3_kernel — reads the pre-LayerNorm tensor x2 saved by the forward kernel, recomputes
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
// Kernel (1): one 128-row tile at a time, fp16 operands, 16 epilogue warps + one MMA-issue warp.
//   * Every operand is fp16 (11-bit mantissa instead of bf16's 8): activations y, h and the weights are O(1); the
//     gradients (grad_out, gPre) are multiplied by a power of two derived from grl_absmax(grad_out) so that
//     max |g| lies in [32, 64) (grl_tc.cuh grad_scale_from_amax) and the factor is removed exactly in the epilogues.
//     tcgen05 kind::f16 needs A and B in ONE format (fp16 x bf16 is an illegal instruction on B200), which is why
//     the gradients cannot simply stay bf16 next to fp16 weights.
//   * b1 rides in the contraction (K = 80: y[:, 64:66] = 1, W1[:, 64:66] = fp16 (hi, lo) split of b1) and the same
//     ones-columns turn the dW1 MMA (N = 80) into the b1 gradient: no bias adds, no 64-value column butterfly.
//   * GELU and GELU' are evaluated two elements per instruction in packed fp16 (gelu_h2); h and dG stay packed.
//   * no fibre recompute: the forward kernel saves the pre-LayerNorm tensor x2 (GrlConvDesc.x2) and this kernel
//     streams it back with cp.async, pulled into L2 one tile earlier still.
// ---------------------------------------------------------------------------------------------------
constexpr int kKb = 80;     // GEMM1 contraction / dW1 width: 64 channels + 2 ones columns + 14 zero columns
constexpr int kLDX2 = 68;   // X2 row stride (floats): conflict-free for the row-per-lane LayerNorm reads

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

// ---------------------------------------------------------------------------------------------------
// Structure: warp-specialised, quarter-pipelined.
//   * 16 epilogue warps (512 threads, thread = tile row x 16-column group) + ONE MMA-issue warp (one elected lane).
//     The roles meet only through mbarriers: `full[b]` (tcgen05.commit -> epilogue, accumulator buffer b holds a fresh
//     result), `edone[b]` (16 warp arrivals -> MMA warp: buffer b has been read and the operand quarter derived from it
//     is in shared memory), `lnready` (y / grad_out operands of the tile are staged), `gyfull` (every MMA of the tile
//     has retired).  No CTA-wide barrier inside the hidden-layer phase.
//   * the hidden layer is walked in four 64-column quarters over TWO 64-column accumulator buffers: while the epilogue
//     warps turn quarter q into h / GELU' (E1) or gPre (E2), the tensor pipe already computes the next quarter into
//     the other buffer, and the weight-gradient / gY MMAs of a finished half run underneath the following epilogues
//     (the tensor pipe retires MMAs in issue order, so a later `full` wait covers them).
//     issue order per tile:  pre0 pre1 | gH0 | gH1 dW2(h0) | pre2 | pre3 gY(h0) dW1(h0) | gH2 | gH3 dW2(h1) | - |
//                            gY(h1) dW1(h1) -> gyfull
//     epilogue order:        E1(0) E1(1) E2(0) E2(1) E1(2) E1(3) E2(2) E2(3), LayerNorm backward, LayerNorm forward
//                            of the next tile.
// TMEM columns: D0 0..63 | D1 64..127 | gY 128..191 | dW1 192..351 (2 x 80) | dW2^T 352..479 (2 x 64).
// Shared memory, operand images, fp16 formats, gradient scale: as kernel (1).
// ---------------------------------------------------------------------------------------------------
constexpr int kNB3Epi = 512;
constexpr int kNB3Threads = kNB3Epi + 32;
constexpr uint32_t kCol3D0 = 0, kCol3D1 = 64, kCol3GY = 128, kCol3DW1 = 192, kCol3DW2 = 352;

struct NodeBwd3Smem {
  __half W1h[kH * kKb];
  __half W2h[kC * kH];
  __half A1[kTM * kKb];
  __half GZh[kTM * kC];
  union {
    struct {
      float X2[kTM * kLDX2];
      float GZf[kTM * kLDX2];
    } in;
    struct {
      __half A2h[kTM * 128];
      __half AP[kTM * 128];
    } h;
  } u;
  float bias[kC], lng[kC], lnb[kC];
  float rs[2][4][kTM];
  float acc_gb2[4][kC], acc_glng[4][kC], acc_glnb[4][kC];
  uint64_t full[2], edone[2], lnready, gyfull, dwdone;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kNB3Threads, 1) fbconv_node_bwd_tc3_kernel(const GrlConvDesc d) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  NodeBwd3Smem& s = *reinterpret_cast<NodeBwd3Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_epi = warp < 16;
  const int q = warp & 3, cg = (warp >> 2) & 3;  // TMEM lane quarter / 16-column group of an epilogue warp
  const int row = 32 * q + lane;

  if (tid == 0) {
    tc::mbar_init(&s.full[0], 1);
    tc::mbar_init(&s.full[1], 1);
    tc::mbar_init(&s.edone[0], 16);
    tc::mbar_init(&s.edone[1], 16);
    tc::mbar_init(&s.lnready, 16);
    tc::mbar_init(&s.gyfull, 1);
    tc::mbar_init(&s.dwdone, 1);
    tc::fence_mbar_init();
  }
  if (warp == 16) tc::tmem_alloc(&s.tmem_base, 512);
  tc::stage_weight_f16(s.W1h, d.w1, kH, kC, kC);
  tc::stage_weight_f16(s.W2h, d.w2, kC, kH, kH);
  for (int n = tid; n < kH; n += kNB3Threads) {
    const float b = d.b1[n];
    const __half hi = __float2half_rn(b);
    const __half lo = __float2half_rn(b - __half2float(hi));
    const __half2 p0 = __halves2half2(hi, lo);
    *reinterpret_cast<uint4*>(s.W1h + ((size_t)8 * kH + n) * 8) = make_uint4(*reinterpret_cast<const uint32_t*>(&p0), 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(s.W1h + ((size_t)9 * kH + n) * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid < kTM) {
    *reinterpret_cast<uint4*>(s.A1 + ((size_t)8 * kTM + tid) * 8) = make_uint4(0x3C003C00u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(s.A1 + ((size_t)9 * kTM + tid) * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid < kC) { s.bias[tid] = d.bias[tid]; s.lng[tid] = d.ln_g[tid]; s.lnb[tid] = d.ln_b[tid]; }
  for (int i = tid; i < 4 * kC; i += kNB3Threads) {
    (&s.acc_gb2[0][0])[i] = 0.f; (&s.acc_glng[0][0])[i] = 0.f; (&s.acc_glnb[0][0])[i] = 0.f;
  }
  const float gscale = d.grad_amax ? tc::grad_scale_from_amax(__ldg(d.grad_amax)) : 1.0f;
  const float inv_gscale = 1.0f / gscale;
  const int n_tiles = (d.n_dst + kTE - 1) / kTE;
  int tile = blockIdx.x;
  auto stage_x2 = [&](int t) {  // epilogue threads only
    const int cnt = min(kTE, d.n_dst - t * kTE);
    const float* src = d.x2 + (size_t)t * kTE * kRow;
    const float* gsrc = d.grad_out + (size_t)t * kTE * kRow;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int f = tid + kNB3Epi * i;
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
  if (is_epi && tile < n_tiles) {
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
  bool first_tile = true;

  if (!is_epi) {
    // ================= MMA-issue warp: warp-uniform loop, one elected lane issues =================
    uint32_t p_ln = 0, p_e0 = 0, p_e1 = 0;
    constexpr uint32_t id_pre = tc::idesc_f16_ex(128, 64, 0, 0, 0, 0);
    constexpr uint32_t id_gh = tc::idesc_f16_ex(128, 64, 0, 1, 0, 0);
    constexpr uint32_t id_dw2 = tc::idesc_f16_ex(128, 64, 1, 1, 0, 0);
    constexpr uint32_t id_dw1 = tc::idesc_f16_ex(128, kKb, 1, 1, 0, 0);
    const tc::Op a1_k = tc::op_k(a1, kTM), a1_mn = tc::op_mn(a1, kTM);
    const tc::Op gz_k = tc::op_k(gz, kTM), gz_mn = tc::op_mn(gz, kTM);
    const tc::Op a2_mn = tc::op_mn(a2, kTM), ap_k = tc::op_k(ap, kTM), ap_mn = tc::op_mn(ap, kTM);
    const tc::Op w1_k = tc::op_k(w1, kH), w1_mn = tc::op_mn(w1, kH), w2_mn = tc::op_mn(w2, kC);
    auto shifted = [](tc::Op o, uint32_t bytes) { o.lo += bytes >> 4; return o; };
    // pre(qq) = [y | 1 1 0..] [W1 | b1]^T, rows 64 qq .. of W1;  gH'(qq) = gZ' W2[:, quarter]  (W2h MN-major: N = k', K = n)
    auto pre = [&](int qq, uint32_t col) { tc::issue_mma_fast<kKb / 16>(tmem + col, a1_k, shifted(w1_k, 64u * qq * 16u), id_pre, false); };
    auto gh = [&](int qq, uint32_t col) { tc::issue_mma_fast<kC / 16>(tmem + col, gz_k, shifted(w2_mn, 8u * qq * (kC * 16u)), id_gh, false); };
    for (; tile < n_tiles; tile += gridDim.x) {
      tc::mbar_wait(&s.lnready, p_ln);
      p_ln ^= 1u;
      tc::tc_fence_after();
      if (tc::elect_one()) {
        pre(0, kCol3D0);
        tc::mma_commit(&s.full[0]);
        pre(1, kCol3D1);
        tc::mma_commit(&s.full[1]);
      }
      __syncwarp();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        tc::mbar_wait(&s.edone[0], p_e0);  // E1(2 hh): h quarter staged, D0 read
        p_e0 ^= 1u;
        tc::tc_fence_after();
        if (tc::elect_one()) {
          gh(2 * hh, kCol3D0);
          tc::mma_commit(&s.full[0]);
        }
        __syncwarp();
        tc::mbar_wait(&s.edone[1], p_e1);  // E1(2 hh + 1)
        p_e1 ^= 1u;
        tc::tc_fence_after();
        if (tc::elect_one()) {
          gh(2 * hh + 1, kCol3D1);
          tc::mma_commit(&s.full[1]);
          // dW2'^T[k'][n] += h^T gZ'  (both MN-major, K = tile rows); retires under the next epilogues
          tc::issue_mma_fast<kTM / 16>(tmem + kCol3DW2 + 64 * hh, a2_mn, gz_mn, id_dw2, !first_tile);
        }
        __syncwarp();
        tc::mbar_wait(&s.edone[0], p_e0);  // E2(2 hh): gPre quarter staged, D0 read
        p_e0 ^= 1u;
        tc::tc_fence_after();
        if (hh == 0) {
          if (tc::elect_one()) {
            pre(2, kCol3D0);
            tc::mma_commit(&s.full[0]);
          }
          __syncwarp();
        }
        tc::mbar_wait(&s.edone[1], p_e1);  // E2(2 hh + 1)
        p_e1 ^= 1u;
        tc::tc_fence_after();
        if (tc::elect_one()) {
          if (hh == 0) {
            pre(3, kCol3D1);
            tc::mma_commit(&s.full[1]);
          }
          // gY' (+)= gPre' W1[half]  (W1h read MN-major: N = c, K = k');  [dW1' | gb1'] += gPre'^T [y | 1 1 0..]
          tc::issue_mma_fast<128 / 16>(tmem + kCol3GY, ap_k, shifted(w1_mn, 128u * hh * 16u), id_gh, hh > 0);
          if (hh == 1) tc::mma_commit(&s.gyfull);  // gY is final: the LayerNorm backward starts under the last dW1 MMAs
          tc::issue_mma_fast<kTM / 16>(tmem + kCol3DW1 + kKb * hh, ap_mn, a1_mn, id_dw1, !first_tile);
          if (hh == 1) tc::mma_commit(&s.dwdone);  // every operand buffer of the tile is free
        }
        __syncwarp();
      }
      first_tile = false;
    }
  } else {
    // ================= epilogue warps =================
    uint32_t pf0 = 0, pf1 = 0, pgy = 0;
    for (; tile < n_tiles; tile += gridDim.x) {
      const int n0 = tile * kTE;
      const int node = n0 + (row >> 4);
      const bool live = node < d.n_dst;
      const size_t roff = (size_t)(live ? node : 0) * kRow + (row & 15) * kC + 16 * cg;
      if (tid == 0) {
        const int nt = tile + gridDim.x;
        if (nt < n_tiles) {
          const uint32_t bytes = (uint32_t)min(kTE, d.n_dst - nt * kTE) * kRow * 4u;
          tc::prefetch_l2(d.x2 + (size_t)nt * kTE * kRow, bytes);
          tc::prefetch_l2(d.grad_out + (size_t)nt * kTE * kRow, bytes);
        }
      }
      cp_async_wait_all();
      tc::group_sync(5, kNB3Epi);

      // ---- LayerNorm forward + scaled grad_out -> fp16 operands ------------------------------------------
      float xh[16];
      float rstd;
      {
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v = ld4(s.u.in.X2 + row * kLDX2 + 16 * cg + 4 * i);
          xh[4 * i] = v.x; xh[4 * i + 1] = v.y; xh[4 * i + 2] = v.z; xh[4 * i + 3] = v.w;
          sum += (v.x + v.y) + (v.z + v.w);
        }
        float g16[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 gv = ld4(s.u.in.GZf + row * kLDX2 + 16 * cg + 4 * i);
          g16[4 * i] = gv.x; g16[4 * i + 1] = gv.y; g16[4 * i + 2] = gv.z; g16[4 * i + 3] = gv.w;
        }
        s.rs[0][cg][row] = sum;
        tc::group_sync(1 + q, 128);
        const float mean = ((s.rs[0][0][row] + s.rs[0][1][row]) + (s.rs[0][2][row] + s.rs[0][3][row])) * (1.0f / 64.0f);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) { xh[i] -= mean; sq = fmaf(xh[i], xh[i], sq); }
        s.rs[1][cg][row] = sq;
        {
          float gs[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) gs[i] = g16[i] * gscale;
          *reinterpret_cast<uint4*>(s.GZh + ((size_t)(2 * cg) * kTM + row) * 8) = tc::pack8_h(gs);
          *reinterpret_cast<uint4*>(s.GZh + ((size_t)(2 * cg + 1) * kTM + row) * 8) = tc::pack8_h(gs + 8);
        }
        warp_colsum16(g16, lane);
        if ((lane & 1) == 0) s.acc_gb2[q][16 * cg + (lane >> 1)] += g16[0];
        tc::group_sync(1 + q, 128);
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
      }
      tc::fence_async_smem();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&s.lnready);

      // ---- hidden layer: four quarters over two accumulator buffers ----------------------------------------
      __half2 dG0[8], dG1[8];
      auto e1 = [&](uint32_t col, uint64_t* fullb, uint32_t& pf, uint64_t* doneb, __half2 (&dG)[8], int chunk) {
        tc::mbar_wait(fullb, pf);
        pf ^= 1u;
        tc::tc_fence_after();
        float v[16];
        tc::tmem_ld16(lane_addr + col + 16 * cg, v);
        __half2 h0[4], h1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          tc::gelu_h2(__floats2half2_rn(v[2 * e], v[2 * e + 1]), h0[e], dG[e]);
          tc::gelu_h2(__floats2half2_rn(v[8 + 2 * e], v[8 + 2 * e + 1]), h1[e], dG[4 + e]);
        }
        *reinterpret_cast<uint4*>(s.u.h.A2h + ((size_t)chunk * kTM + row) * 8) = tc::pack_h8(h0);
        *reinterpret_cast<uint4*>(s.u.h.A2h + ((size_t)(chunk + 1) * kTM + row) * 8) = tc::pack_h8(h1);
        tc::fence_async_smem();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(doneb);
      };
      auto e2 = [&](uint32_t col, uint64_t* fullb, uint32_t& pf, uint64_t* doneb, const __half2 (&dG)[8], int chunk) {
        tc::mbar_wait(fullb, pf);
        pf ^= 1u;
        tc::tc_fence_after();
        float v[16];
        tc::tmem_ld16(lane_addr + col + 16 * cg, v);
        __half2 p0[4], p1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          p0[e] = __hmul2(__floats2half2_rn(v[2 * e], v[2 * e + 1]), dG[e]);
          p1[e] = __hmul2(__floats2half2_rn(v[8 + 2 * e], v[8 + 2 * e + 1]), dG[4 + e]);
        }
        *reinterpret_cast<uint4*>(s.u.h.AP + ((size_t)chunk * kTM + row) * 8) = tc::pack_h8(p0);
        *reinterpret_cast<uint4*>(s.u.h.AP + ((size_t)(chunk + 1) * kTM + row) * 8) = tc::pack_h8(p1);
        tc::fence_async_smem();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(doneb);
      };
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        e1(kCol3D0, &s.full[0], pf0, &s.edone[0], dG0, 2 * cg);
        e1(kCol3D1, &s.full[1], pf1, &s.edone[1], dG1, 8 + 2 * cg);
        e2(kCol3D0, &s.full[0], pf0, &s.edone[0], dG0, 2 * cg);
        e2(kCol3D1, &s.full[1], pf1, &s.edone[1], dG1, 8 + 2 * cg);
      }
      tc::mbar_wait(&s.gyfull, pgy);  // gY is final
      tc::tc_fence_after();

      // ---- LayerNorm backward: gY (TMEM) -> g_x2 (HBM), g_ln_g, g_ln_b -----------------------------------
      {
        float gy[32];
        {
          float v[16];
          tc::tmem_ld16(lane_addr + kCol3GY + 16 * cg, v);
#pragma unroll
          for (int e = 0; e < 16; ++e) gy[e] = v[e] * inv_gscale;
        }
        tc::mbar_wait(&s.dwdone, pgy);  // the last dW1 MMAs have retired: A2h / AP (= the staging buffers) and A1 are free
        pgy ^= 1u;
        {
          const int nt = tile + gridDim.x;
          if (nt < n_tiles) {
            stage_x2(nt);
            cp_async_commit();
          }
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
        tc::group_sync(1 + q, 128);
        const float m1 = ((s.rs[0][0][row] + s.rs[0][1][row]) + (s.rs[0][2][row] + s.rs[0][3][row])) * (1.0f / 64.0f);
        const float m2 = ((s.rs[1][0][row] + s.rs[1][1][row]) + (s.rs[1][2][row] + s.rs[1][3][row])) * (1.0f / 64.0f);
        float* gdst = d.grad_x2 + roff;
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
#pragma unroll
        for (int i = 0; i < 16; ++i) gy[16 + i] = gy[i] * xh[i];
        tc::warp_colsum<32>(gy, lane);
        if (lane < 16) s.acc_glnb[q][16 * cg + lane] += gy[0];
        else s.acc_glng[q][16 * cg + lane - 16] += gy[0];
      }
      tc::tc_fence_before();
      first_tile = false;
    }
    cp_async_wait_all();
  }

  // ---- write this CTA's partial slot ---------------------------------------------------------------------------
  __syncthreads();
  tc::tc_fence_after();
  float* P = d.node_grad_partials + (size_t)blockIdx.x * GRL_NODE_GRAD_FLOATS;
  if (is_epi) {
#pragma unroll 1
    for (int h2 = 0; h2 < 2; ++h2) {
      const int c0 = 16 * cg;
      float v[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] = 0.f;
      if (!first_tile) tc::tmem_ld16(lane_addr + kCol3DW1 + kKb * h2 + c0, v);
      float* p1 = P + kPGW1 + (size_t)(128 * h2 + row) * kC + c0;
#pragma unroll
      for (int e = 0; e < 16; e += 4)
        st4(p1 + e, make_float4(v[e] * inv_gscale, v[e + 1] * inv_gscale, v[e + 2] * inv_gscale, v[e + 3] * inv_gscale));
      if (cg == 0) {
        float b[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) b[e] = 0.f;
        if (!first_tile) tc::tmem_ld16(lane_addr + kCol3DW1 + kKb * h2 + 64, b);
        P[kPGB1 + 128 * h2 + row] = b[0] * inv_gscale;
      }
      if (!first_tile) tc::tmem_ld16(lane_addr + kCol3DW2 + 64 * h2 + c0, v);
#pragma unroll
      for (int e = 0; e < 16; ++e) P[kPGW2 + (size_t)(c0 + e) * kH + 128 * h2 + row] = v[e] * inv_gscale;
    }
  }
  if (tid < kC) {
    P[kPGB2 + tid] = ((s.acc_gb2[0][tid] + s.acc_gb2[1][tid]) + s.acc_gb2[2][tid]) + s.acc_gb2[3][tid];
    P[kPGLNG + tid] = ((s.acc_glng[0][tid] + s.acc_glng[1][tid]) + s.acc_glng[2][tid]) + s.acc_glng[3][tid];
    P[kPGLNB + tid] = ((s.acc_glnb[0][tid] + s.acc_glnb[1][tid]) + s.acc_glnb[2][tid]) + s.acc_glnb[3][tid];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 16) tc::tmem_dealloc(tmem, 512);
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
  // both orientations of the thread ride in one packed register pair: (fk[2 op][p], fk[2 op + 1][p]) etc.
  unsigned long long fk2[kO], gfk2[kO];
#pragma unroll
  for (int p = 0; p < kO; ++p) {
    fk2[p] = pack2(__ldg(d.fiber_kernel + ((size_t)((2 * op) * kO + p)) * kC + c) * 0.0625f,
                   __ldg(d.fiber_kernel + ((size_t)((2 * op + 1) * kO + p)) * kC + c) * 0.0625f);
    gfk2[p] = pack2(0.f, 0.f);
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
      const unsigned long long xv = pack2(X[(2 * op) * kC], X[(2 * op + 1) * kC]);
      unsigned long long a2 = pack2(0.f, 0.f);
      float gsum = 0.f;
#pragma unroll
      for (int p = 0; p < kO; ++p) {
        const float g = G[p * kC];
        const unsigned long long gg = pack2(g, g);
        a2 = ffma2(gg, fk2[p], a2);          // g_x1[2 op + oi] += g_x2[p] fk[2 op + oi][p]
        gfk2[p] = ffma2(xv, gg, gfk2[p]);    // g_fk[2 op + oi][p] += x1[2 op + oi] g_x2[p]
        gsum += g;
      }
      if (node < d.n_dst) {
        float a0, a1;
        unpack2(a2, a0, a1);
        d.grad_x1[(size_t)node * kRow + (2 * op) * kC + c] = a0;
        d.grad_x1[(size_t)node * kRow + (2 * op + 1) * kC + c] = a1;
      }
      if (op == 0) gbias += gsum;
    }
    __syncthreads();  // everyone done with `buf` before the next iteration's prefetch overwrites it
  }
  float* P = d.node_grad_partials + (size_t)blockIdx.x * GRL_NODE_GRAD_FLOATS;
#pragma unroll
  for (int p = 0; p < kO; ++p) {
    float g0, g1;
    unpack2(gfk2[p], g0, g1);
    P[kPGFK + ((size_t)((2 * op) * kO + p)) * kC + c] = g0 * 0.0625f;
    P[kPGFK + ((size_t)((2 * op + 1) * kO + p)) * kC + c] = g1 * 0.0625f;
  }
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
    const int smem = (int)sizeof(grl::NodeBwd3Smem);
    if (grl::ensure_dynamic_smem((const void*)grl::fbconv_node_bwd_tc3_kernel, smem) != GRL_OK) return GRL_ECUDA;
    grl::fbconv_node_bwd_tc3_kernel<<<d->n_partials_node, grl::kNB3Threads, smem, (cudaStream_t)stream>>>(*d);
  }
  int rc = grl::check_launch("grl_fbconv_node_bwd_tc (mlp)");
  if (rc != GRL_OK) return rc;
  const int smem2 = (int)sizeof(grl::FiberBwdSmem);
  if (grl::ensure_dynamic_smem((const void*)grl::fbconv_fiber_bwd_kernel, smem2) != GRL_OK) return GRL_ECUDA;
  grl::fbconv_fiber_bwd_kernel<<<d->n_partials_node, grl::kFiberThreads, smem2, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_fbconv_node_bwd_tc (fibre)");
}
