// Node half of the fibre-bundle convolution, tensor-core path (tcgen05.mma, accumulators in TMEM).
//   x2 = fibre(x1) + bias;  y = LayerNorm(x2);  h = GELU(y W1^T + b1);  out = x_dst + h W2^T + b2
// Same math and reference lines as grl_conv_node.cu (ponita/conv.py:88-114, ponita/ponita.py:163-175,219-230);
// the two contractions run as [128 x 80] x [80 x 256] and [128 x 256] x [256 x 64] fp16 MMAs with fp32
// accumulation, everything else (fibre conv, LayerNorm, residual) stays fp32 on the CUDA cores.
#include <cuda.h>

#include "grl_common.cuh"
#include "grl_tc.cuh"

namespace grl {

constexpr int kLDX = 72;  // X2 row stride (floats): conflict-free for the thread-pair-per-row LayerNorm reads

// ---------------------------------------------------------------------------------------------------
// Two independent tile pipelines per SM.
// A single 256-thread pipeline is issue / latency bound (ncu r01: smsp__issue_active 33 %, 8 warps per SM, every
// phase fenced by a CTA barrier).  Here one 512-thread CTA per SM runs TWO 256-thread groups that walk alternate tiles with
// their own buffers, mbarriers, TMEM columns and named barriers, sharing only the weight images, so one group's
// MMA / TMEM / global-memory waits are covered by the other group's CUDA-core phases.  Further changes:
//   * operands are fp16 (LayerNorm and GELU outputs are O(1); 11-bit mantissa instead of bf16's 8),
//   * b1 rides in the contraction: K = 80 with y[:, 64:66] = 1 and W1[:, 64:66] = (hi, lo) fp16 split of b1,
//   * GELU is evaluated two elements per instruction (HFMA2 / MUFU.TANH.F16, grl_tc.cuh gelu_h2),
//   * the hidden activations never touch shared memory: each thread packs GELU(D1) of its row into fp16 pairs and stores
//     them back into its own TMEM lane (tcgen05.st) over the D1 columns it has already read; GEMM2 takes that as its A
//     operand (TS mode) and accumulates into D1 columns 64..127, so a group still needs 256 columns;
//   * with no hidden-layer image aliasing the x1 tile, the next tile's x1 rows are requested as soon as GEMM1 has retired;
//   * the residual rows x_dst of the tile arrive through tensor-map TMA (two 128-row x 32-channel boxes, 128-byte swizzled)
//     into the bytes of X2 once the LayerNorm has consumed them - no residual values parked in registers across GEMM2
//     (at the 128-register cap they went through local memory and their load latency sat in front of the epilogue).
// ---------------------------------------------------------------------------------------------------
constexpr int kNF2Threads = 512;
constexpr int kK1 = 80;  // GEMM1 contraction length: 64 channels + 2 bias columns + 14 zero columns

struct NodeFwd2Group {
  union {
    struct {
      float X1[kTM * kC];    // x1 tile, dense rows (phase A reads one row per warp: conflict-free)
      float X2[kTM * kLDX];  // fibre conv output (phase B reads thread-pair-per-row: stride 72)
    } x;
    __half A1[kTM * kK1];    // y: [10 chunks][128 rows][8]   (aliases X1, dead after phase A)
  } u;
};

struct NodeFwd2Smem {
  __half W1h[kH * kK1];      // [10 chunks][256 rows][8]; chunk 8 = (b1 hi, b1 lo, 0 ...), chunk 9 = 0
  __half W2h[kC * kH];       // [32 chunks][64 rows][8]
  NodeFwd2Group g[2];
  float b2[kC], bias[kC], lng[kC], lnb[kC];
  uint64_t bar[2][2];
  uint64_t bar_x[2];         // per group: transaction barrier of the bulk copy that stages its x1 tile
  uint64_t bar_d[2];         // per group: transaction barrier of the TMA boxes that stage its x_dst tile
  uint32_t tmem_base;
};

// x1 tile ([cnt <= 8 nodes][16][64] fp32, contiguous in HBM) -> dense X1: ONE bulk copy of cnt * 4 KB issued by the
// group's first thread (completion counted in bytes on `bar`), rows of missing nodes zero-filled by everybody.
__device__ __forceinline__ void stage_x1_dense(float* __restrict__ X1, const float* __restrict__ src, int cnt, int gt,
                                               uint64_t* bar) {
  if (gt == 0) {
    tc::mbar_expect_tx(bar, (uint32_t)cnt * kRow * 4u);
    tc::bulk_g2s(X1, src, (uint32_t)cnt * kRow * 4u, bar);
  }
  if (cnt < kTE) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = gt + 256 * i;  // float4 index 0..2047 of the [128][64] tile
      if ((f >> 8) >= cnt) *reinterpret_cast<float4*>(X1 + 4 * f) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

template <bool kAcc>
__global__ void __launch_bounds__(kNF2Threads, 1) fbconv_node_fwd_tc2_kernel(const GrlConvDesc d, const __grid_constant__ CUtensorMap tm_xd) {
  extern __shared__ unsigned char smem_raw[];
  NodeFwd2Smem& s = *reinterpret_cast<NodeFwd2Smem*>(
      smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));  // the swizzle atoms want a 1024-byte aligned base
  const int tid = threadIdx.x;
  const int grp = tid >> 8, gt = tid & 255, gw = gt >> 5, lane = tid & 31;
  NodeFwd2Group& G = s.g[grp];
  const int bar_id = 1 + grp;

  // ---- one-time setup (whole CTA) --------------------------------------------------------------------
  if (tid == 0) {
    tc::mbar_init(&s.bar[0][0], 1);
    tc::mbar_init(&s.bar[0][1], 1);
    tc::mbar_init(&s.bar[1][0], 1);
    tc::mbar_init(&s.bar[1][1], 1);
    tc::mbar_init(&s.bar_x[0], 1);
    tc::mbar_init(&s.bar_x[1], 1);
    tc::mbar_init(&s.bar_d[0], 1);
    tc::mbar_init(&s.bar_d[1], 1);
    tc::fence_mbar_init();
  }
  if (tid < 32) tc::tmem_alloc(&s.tmem_base, 512);
  tc::stage_weight_f16(s.W1h, d.w1, kH, kC, kC);   // chunks 0..7 of W1h
  tc::stage_weight_f16(s.W2h, d.w2, kC, kH, kH);
  for (int n = tid; n < kH; n += kNF2Threads) {    // chunks 8 (bias split) and 9 (zero)
    const float b = d.b1[n];
    const __half hi = __float2half_rn(b);
    const __half lo = __float2half_rn(b - __half2float(hi));
    const __half z = __float2half_rn(0.f);
    const __half2 h[4] = {__halves2half2(hi, lo), __halves2half2(z, z), __halves2half2(z, z), __halves2half2(z, z)};
    *reinterpret_cast<uint4*>(s.W1h + ((size_t)8 * kH + n) * 8) = tc::pack_h8(h);
    *reinterpret_cast<uint4*>(s.W1h + ((size_t)9 * kH + n) * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid < kC) { s.b2[tid] = d.b2[tid]; s.bias[tid] = d.bias[tid]; s.lng[tid] = d.ln_g[tid]; s.lnb[tid] = d.ln_b[tid]; }
  // fibre kernel slice of this thread: channel fc, output orientations p = 4 pq .. 4 pq + 3, all 16 inputs o
  const int fc = gt & 63, pq = gt >> 6;
  // packed pairs of output orientations (FFMA2: two fused multiply-adds per fma-pipe issue slot)
  unsigned long long fk2[kO][2];
#pragma unroll
  for (int o = 0; o < kO; ++o)
#pragma unroll
    for (int pi = 0; pi < 2; ++pi)
      fk2[o][pi] = pack2(__ldg(d.fiber_kernel + ((size_t)(o * kO + 4 * pq + 2 * pi)) * kC + fc) * 0.0625f,
                         __ldg(d.fiber_kernel + ((size_t)(o * kO + 4 * pq + 2 * pi + 1)) * kC + fc) * 0.0625f);

  const int n_tiles = (d.n_dst + kTE - 1) / kTE;
  const int tile_stride = 2 * gridDim.x;
  int tile = 2 * blockIdx.x + grp;
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();  // weight images, mbarriers and the TMEM base are visible to everybody
  tc::tc_fence_after();
  if (tile < n_tiles)
    stage_x1_dense(G.u.x.X1, d.x1 + (size_t)tile * kTE * kRow, min(kTE, d.n_dst - tile * kTE), gt, &s.bar_x[grp]);
  const uint32_t tmem = s.tmem_base + 256u * grp;
  const uint32_t a1_addr = tc::smem_u32(G.u.A1);
  const uint32_t w1_addr = tc::smem_u32(s.W1h), w2_addr = tc::smem_u32(s.W2h);
  const int q = gw & 3, hh = gw >> 2;
  const int row = 32 * q + lane;
  const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
  uint32_t parity = 0;

  for (; tile < n_tiles; tile += tile_stride) {
    const int n0 = tile * kTE;
    if (gt == 0) {  // L2 prefetch: this tile's residual rows (read before the GEMM2 wait) and the group's next x1 tile
      tc::prefetch_l2(d.x_dst + (size_t)n0 * kRow, (uint32_t)min(kTE, d.n_dst - n0) * kRow * 4u);
      const int nt = tile + tile_stride;
      if (nt < n_tiles) tc::prefetch_l2(d.x1 + (size_t)nt * kTE * kRow, (uint32_t)min(kTE, d.n_dst - nt * kTE) * kRow * 4u);
    }
    tc::mbar_wait(&s.bar_x[grp], parity);  // the bulk copy of this tile's x1 rows has landed
    tc::group_sync(bar_id, 256);           // (and the zero rows of a partial tile are visible to the group)

    // ---- A: fibre convolution + bias -------------------------------------------------------------
    {
      const float bias_c = s.bias[fc];
#pragma unroll 2
      for (int j = 0; j < kTE; ++j) {
        unsigned long long a01 = pack2(0.f, 0.f), a23 = a01;
#pragma unroll
        for (int o = 0; o < kO; ++o) {
          const float x = G.u.x.X1[(16 * j + o) * kC + fc];
          const unsigned long long xx = pack2(x, x);
          a01 = ffma2(xx, fk2[o][0], a01);
          a23 = ffma2(xx, fk2[o][1], a23);
        }
        float a0, a1, a2, a3;
        unpack2(a01, a0, a1);
        unpack2(a23, a2, a3);
        float* o2 = G.u.x.X2 + (16 * j + 4 * pq) * kLDX + fc;
        o2[0 * kLDX] = a0 + bias_c;
        o2[1 * kLDX] = a1 + bias_c;
        o2[2 * kLDX] = a2 + bias_c;
        o2[3 * kLDX] = a3 + bias_c;
      }
    }
    tc::group_sync(bar_id, 256);

    // ---- B: LayerNorm over the 64 channels of a row; thread pair (row r, half h) -> fp16 operand A1 ------
    {
      const int r = gt >> 1, h = gt & 1;
      float4 v[8];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i] = ld4(G.u.x.X2 + r * kLDX + 4 * (2 * i + h));
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
      if (d.x2 && n0 + (r >> 4) < d.n_dst) {  // saved for the backward pass (it then skips the fibre recompute)
        float* xo = d.x2 + (size_t)(n0 + (r >> 4)) * kRow + (r & 15) * kC + 4 * h;
#pragma unroll
        for (int i = 0; i < 8; ++i) st4(xo + 8 * i, v[i]);
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      const float mean = sum * (1.0f / 64.0f);
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
      }
      sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      const float rstd = rsqrtf(sq * (1.0f / 64.0f) + 1e-5f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = 4 * (2 * i + h);  // channels c..c+3 = half `h` of 16-byte chunk i
        const float4 g = ld4(s.lng + c), b = ld4(s.lnb + c);
        const __half2 p0 = __floats2half2_rn(v[i].x * rstd * g.x + b.x, v[i].y * rstd * g.y + b.y);
        const __half2 p1 = __floats2half2_rn(v[i].z * rstd * g.z + b.z, v[i].w * rstd * g.w + b.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&p0);
        pk.y = *reinterpret_cast<const uint32_t*>(&p1);
        *reinterpret_cast<uint2*>(G.u.A1 + ((size_t)i * kTM + r) * 8 + 4 * h) = pk;
      }
      // chunk 8 = (1, 1, 0 ...) multiplies the (hi, lo) bias columns of W1h; chunk 9 = 0
      *reinterpret_cast<uint4*>(G.u.A1 + ((size_t)(8 + h) * kTM + r) * 8) = make_uint4(h == 0 ? 0x3C003C00u : 0u, 0u, 0u, 0u);
    }
    // ---- C: GEMM1  D1[128 x 256] = [y | 1 1 0..] [W1 | b1_hi b1_lo 0..]^T ------------------------------
    tc::fence_async_smem();
    tc::tc_fence_before();
    tc::group_sync(bar_id, 256);
    if (gt == 0) {
      tc::tc_fence_after();
      tc::issue_mma_rolled(tmem, tc::view_k(a1_addr, kTM), tc::view_k(w1_addr, kH), tc::idesc_f16_ex(128, kH, 0, 0, 0, 0), kK1 / 16, false);
      tc::mma_commit(&s.bar[grp][0]);
      // X2 has been consumed by the LayerNorm (barrier above): the tile's residual rows land there under GEMM1 .. GEMM2.
      // Rows past the end of the tensor are zero-filled by the TMA unit and still count towards the byte total.
      tc::mbar_expect_tx(&s.bar_d[grp], 2u * kTM * 32u * 4u);
      tc::tma_load_2d(G.u.x.X2, &tm_xd, 0, n0 * kO, &s.bar_d[grp]);
      tc::tma_load_2d(G.u.x.X2 + kTM * 32, &tm_xd, 32, n0 * kO, &s.bar_d[grp]);
    }
    // ---- D: hidden = GELU(D1), packed fp16, back into this thread's TMEM lane (A operand of GEMM2) -----------------
    tc::mbar_wait(&s.bar[grp][0], parity);
    tc::tc_fence_after();
    // A1 (= the X1 bytes) is free: request the group's next x1 tile now, it lands under the GELU and GEMM2
    {
      const int nt = tile + tile_stride;
      if (nt < n_tiles) {
        tc::fence_async_smem();  // generic-proxy accesses to the tile before the bulk engine rewrites it
        stage_x1_dense(G.u.x.X1, d.x1 + (size_t)nt * kTE * kRow, min(kTE, d.n_dst - nt * kTE), gt, &s.bar_x[grp]);
      }
    }
    {
#pragma unroll 2
      for (int i = 0; i < 8; ++i) {
        const int c0 = 128 * hh + 16 * i;  // hidden units c0 .. c0 + 15 of this thread's row
        float v[16];
        tc::tmem_ld16(lane_addr + c0, v);
        uint32_t r[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __half2 a = tc::gelu_h2(__floats2half2_rn(v[2 * e], v[2 * e + 1]));
          const __half2 b = tc::gelu_h2(__floats2half2_rn(v[8 + 2 * e], v[8 + 2 * e + 1]));
          r[e] = *reinterpret_cast<const uint32_t*>(&a);
          r[4 + e] = *reinterpret_cast<const uint32_t*>(&b);
        }
        // packed columns 128 hh + 8 i .. + 7: always behind this thread's own reads, never in the other half's range
        tc::tmem_st8(lane_addr + 128 * hh + 8 * i, r);
      }
      tc::tmem_st_wait();
    }
    // ---- E: GEMM2  D2[128 x 64] = hidden W2^T, A from tensor memory, into D1 columns 64..127 (all read by now) ----
    tc::tc_fence_before();
    tc::group_sync(bar_id, 256);
    if (gt == 0) {
      tc::tc_fence_after();
      const tc::OpView w2v = tc::view_k(w2_addr, kC);
      tc::OpView w2hi = w2v;
      w2hi.addr += 8 * w2v.adv;  // K steps 8..15 (hidden units 128..255)
      tc::issue_mma_ts(tmem + 64, tmem, w2v, tc::idesc_f16_ex(128, kC, 0, 0, 0, 0), 8, false);
      tc::issue_mma_ts(tmem + 64, tmem + 128, w2hi, tc::idesc_f16_ex(128, kC, 0, 0, 0, 0), 8, true);
      tc::mma_commit(&s.bar[grp][1]);
    }
    const int node = n0 + (row >> 4);
    const bool live = node < d.n_dst;
    const size_t off = (size_t)(live ? node : 0) * kRow + (row & 15) * kC + 32 * hh;
    tc::mbar_wait(&s.bar[grp][1], parity);
    tc::tc_fence_after();
    // ---- F: out = x_dst + D2 + b2 (x_dst from the swizzled tile: chunk k of row r sits at chunk position k ^ (r & 7)) ----
    tc::mbar_wait(&s.bar_d[grp], parity);
    const float* xrow = G.u.x.X2 + hh * (kTM * 32) + row * 32;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float v[16];
      tc::tmem_ld16(lane_addr + 64 + 32 * hh + 16 * i, v);
      if (live) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 bb = ld4(s.b2 + 32 * hh + 16 * i + 4 * e);
          float4 x = ld4(xrow + (((4 * i + e) ^ (row & 7)) << 2));
          if (kAcc) {  // HeteroConv group "sum": out already holds the other edge type's result
            const float4 old = ldg4(d.out + off + 16 * i + 4 * e);
            x.x += old.x; x.y += old.y; x.z += old.z; x.w += old.w;
          }
          st4(d.out + off + 16 * i + 4 * e, make_float4(x.x + (v[4 * e] + bb.x), x.y + (v[4 * e + 1] + bb.y),
                                                       x.z + (v[4 * e + 2] + bb.z), x.w + (v[4 * e + 3] + bb.w)));
        }
      }
    }
    tc::tc_fence_before();  // TMEM reads of this tile ordered before the group's next MMAs (after its next barrier)
    parity ^= 1u;
  }
  cp_async_wait_all();
  tc::tc_fence_before();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(s.tmem_base, 512);
}

}  // namespace grl

extern "C" int grl_fbconv_node_fwd_tc(const GrlConvDesc* d, grl_stream_t stream) {
  GRL_REQUIRE(d && d->n_dst > 0, GRL_EINVAL, "grl_fbconv_node_fwd_tc: bad descriptor");
  GRL_REQUIRE(d->x1 && d->fiber_kernel && d->bias && d->ln_g && d->ln_b && d->w1 && d->b1 && d->w2 && d->b2 && d->x_dst &&
                  d->out, GRL_EINVAL, "grl_fbconv_node_fwd_tc: null pointer");
  const int n_tiles = (d->n_dst + grl::kTE - 1) / grl::kTE;
  const int smem2 = (int)sizeof(grl::NodeFwd2Smem) + 1024;
  if (grl::ensure_dynamic_smem((const void*)grl::fbconv_node_fwd_tc2_kernel<false>, smem2) != GRL_OK) return GRL_ECUDA;
  if (grl::ensure_dynamic_smem((const void*)grl::fbconv_node_fwd_tc2_kernel<true>, smem2) != GRL_OK) return GRL_ECUDA;
  alignas(64) CUtensorMap tm_xd;  // x_dst as [n_dst * 16 rows][64]: boxes of a whole 8-node tile x one channel half
  if (grl::make_row_tensor_map(&tm_xd, d->x_dst, d->n_dst, grl::kTM) != GRL_OK) return GRL_ECUDA;
  int grid = grl::sm_count();
  if (2 * grid > n_tiles) grid = (n_tiles + 1) / 2;
  if (d->accumulate_out) grl::fbconv_node_fwd_tc2_kernel<true><<<grid, grl::kNF2Threads, smem2, (cudaStream_t)stream>>>(*d, tm_xd);
  else grl::fbconv_node_fwd_tc2_kernel<false><<<grid, grl::kNF2Threads, smem2, (cudaStream_t)stream>>>(*d, tm_xd);
  return grl::check_launch("grl_fbconv_node_fwd_tc");
}
