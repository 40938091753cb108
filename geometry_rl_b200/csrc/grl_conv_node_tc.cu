// Node half of the fibre-bundle convolution, bf16 tensor-core path (tcgen05.mma, accumulators in TMEM).
//   x2 = fibre(x1) + bias;  y = LayerNorm(x2);  h = GELU(y W1^T + b1);  out = x_dst + h W2^T + b2
// Same math and reference lines as grl_conv_node.cu (ponita/conv.py:88-114, ponita/ponita.py:163-175,219-230);
// the two contractions run as [128 x 64] x [64 x 256] and [128 x 256] x [256 x 64] bf16 MMAs with fp32
// accumulation, everything else (fibre conv, LayerNorm, GELU, residual) stays fp32 on the CUDA cores.
//
// One persistent CTA (256 threads) per SM; per tile of 8 nodes = 128 rows:
//   A  x1 tile (cp.async) -> fibre conv: thread (c, p-quarter) keeps fk[16][4] in registers       -> X2 (smem fp32)
//   B  LayerNorm: thread pair per row                                                            -> A1 (bf16 operand)
//   C  one thread issues GEMM1 (4 x tcgen05.mma, N = 256)                                        -> TMEM cols 0..255
//   D  all warps: tcgen05.ld, + b1, GELU                                                          -> A2 (bf16 operand)
//   E  one thread issues GEMM2 (16 x tcgen05.mma, N = 64)                                        -> TMEM cols 256..319
//   F  all warps: tcgen05.ld, + b2 + x_dst                                                        -> out (HBM)
// The next tile's x1 is prefetched with cp.async while phase F streams the residual / output.
#include "grl_common.cuh"
#include "grl_tc.cuh"

namespace grl {

constexpr int kLDX = 72;  // fp32 tile row stride (floats): conflict-free for both access patterns below

struct NodeTcSmem {
  __nv_bfloat16 W1b[kH * kC];  // B operand of GEMM1: [8 chunks][256 rows][8]
  __nv_bfloat16 W2b[kC * kH];  // B operand of GEMM2: [32 chunks][64 rows][8]
  __nv_bfloat16 A1[kTM * kC];  // y tile: [8 chunks][128 rows][8]
  union {
    struct {
      float X1[kTM * kLDX];
      float X2[kTM * kLDX];
    } x;
    __nv_bfloat16 A2[kTM * kH];  // hidden tile: [32 chunks][128 rows][8]
  } u;
  float b1[kH];
  float b2[kC], bias[kC], lng[kC], lnb[kC];
  uint64_t bar[2];
  uint32_t tmem_base;
};

// x1 rows of `cnt` (<= 8) consecutive nodes -> X1 (stride kLDX), rows of missing nodes zero-filled
__device__ __forceinline__ void stage_x1(float* __restrict__ X1, const float* __restrict__ src, int cnt) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = threadIdx.x + kThreads * i;  // float4 index 0..2047
    const int row = f >> 4, c4 = f & 15;
    float* d = X1 + row * kLDX + 4 * c4;
    if ((row >> 4) < cnt) cp_async16(d, src + (size_t)row * kC + 4 * c4);
    else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(kThreads, 1) fbconv_node_fwd_tc_kernel(const GrlConvDesc d) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  NodeTcSmem& s = *reinterpret_cast<NodeTcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup -----------------------------------------------------------------------------
  if (tid == 0) {
    tc::mbar_init(&s.bar[0], 1);
    tc::mbar_init(&s.bar[1], 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&s.tmem_base, 512);
  tc::stage_weight_bf16(s.W1b, d.w1, kH, kC, kC);   // W1 [256][64]
  tc::stage_weight_bf16(s.W2b, d.w2, kC, kH, kH);   // W2 [64][256]
  for (int i = tid; i < kH; i += kThreads) s.b1[i] = d.b1[i];
  if (tid < kC) { s.b2[tid] = d.b2[tid]; s.bias[tid] = d.bias[tid]; s.lng[tid] = d.ln_g[tid]; s.lnb[tid] = d.ln_b[tid]; }
  // fibre kernel slice of this thread: channel c, output orientations p = 4 pq .. 4 pq + 3, all 16 inputs o
  const int fc = tid & 63, pq = tid >> 6;
  float fk[kO][4];
#pragma unroll
  for (int o = 0; o < kO; ++o)
#pragma unroll
    for (int pi = 0; pi < 4; ++pi) fk[o][pi] = __ldg(d.fiber_kernel + ((size_t)(o * kO + 4 * pq + pi)) * kC + fc) * 0.0625f;

  const int n_tiles = (d.n_dst + kTE - 1) / kTE;
  int tile = blockIdx.x;
  if (tile < n_tiles) {
    stage_x1(s.u.x.X1, d.x1 + (size_t)tile * kTE * kRow, min(kTE, d.n_dst - tile * kTE));
    cp_async_commit();
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = s.tmem_base;
  const uint32_t a1_addr = tc::smem_u32(s.A1), a2_addr = tc::smem_u32(s.u.A2);
  const uint32_t w1_addr = tc::smem_u32(s.W1b), w2_addr = tc::smem_u32(s.W2b);
  uint32_t parity = 0;

  for (; tile < n_tiles; tile += gridDim.x) {
    const int n0 = tile * kTE, cnt = min(kTE, d.n_dst - n0);
    cp_async_wait_all();
    __syncthreads();  // X1 of this tile visible to everyone

    // ---- A: fibre convolution + bias -------------------------------------------------------------
    {
      const float bias_c = s.bias[fc];
#pragma unroll 2
      for (int j = 0; j < kTE; ++j) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int o = 0; o < kO; ++o) {
          const float x = s.u.x.X1[(16 * j + o) * kLDX + fc];
          a0 = fmaf(x, fk[o][0], a0);
          a1 = fmaf(x, fk[o][1], a1);
          a2 = fmaf(x, fk[o][2], a2);
          a3 = fmaf(x, fk[o][3], a3);
        }
        float* o2 = s.u.x.X2 + (16 * j + 4 * pq) * kLDX + fc;
        o2[0 * kLDX] = a0 + bias_c;
        o2[1 * kLDX] = a1 + bias_c;
        o2[2 * kLDX] = a2 + bias_c;
        o2[3 * kLDX] = a3 + bias_c;
      }
    }
    __syncthreads();

    // ---- B: LayerNorm over the 64 channels of a row; thread pair (row r, half h) -----------------------
    {
      const int r = tid >> 1, h = tid & 1;
      float4 v[8];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i] = ld4(s.u.x.X2 + r * kLDX + 4 * (2 * i + h));
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
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
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[i].x * rstd * g.x + b.x, v[i].y * rstd * g.y + b.y);
        __nv_bfloat162 p1 = __floats2bfloat162_rn(v[i].z * rstd * g.z + b.z, v[i].w * rstd * g.w + b.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&p0);
        pk.y = *reinterpret_cast<uint32_t*>(&p1);
        *reinterpret_cast<uint2*>(s.A1 + ((size_t)i * kTM + r) * 8 + 4 * h) = pk;
      }
    }
    // ---- C: GEMM1  D1[128 x 256] = y W1^T ---------------------------------------------------------
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      tc::issue_gemm<kH, kC>(tmem, a1_addr, kTM, w1_addr, kH, false);
      tc::mma_commit(&s.bar[0]);
    }
    // ---- D: hidden = GELU(D1 + b1) -> A2 (aliases X1 / X2, which nobody reads any more) ----------------
    tc::mbar_wait(&s.bar[0], parity);
    tc::tc_fence_after();
    {
      const int q = warp & 3, hh = warp >> 2;
      const int row = 32 * q + lane;
#pragma unroll 1
      for (int i = 0; i < 8; ++i) {
        const int c0 = 128 * hh + 16 * i;
        float v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + c0, v);
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const float4 bb = ld4(s.b1 + c0 + e);
          v[e] = gelu_fast(v[e] + bb.x);
          v[e + 1] = gelu_fast(v[e + 1] + bb.y);
          v[e + 2] = gelu_fast(v[e + 2] + bb.z);
          v[e + 3] = gelu_fast(v[e + 3] + bb.w);
        }
        *reinterpret_cast<uint4*>(s.u.A2 + ((size_t)(c0 >> 3) * kTM + row) * 8) = tc::pack8(v);
        *reinterpret_cast<uint4*>(s.u.A2 + ((size_t)((c0 >> 3) + 1) * kTM + row) * 8) = tc::pack8(v + 8);
      }
    }
    // ---- E: GEMM2  D2[128 x 64] = hidden W2^T ---------------------------------------------------------
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      tc::issue_gemm<kC, kH>(tmem + kH, a2_addr, kTM, w2_addr, kC, false);
      tc::mma_commit(&s.bar[1]);
    }
    tc::mbar_wait(&s.bar[1], parity);
    tc::tc_fence_after();
    // A2 (= X1 / X2) is free again: prefetch the next tile's x1 while the epilogue streams out
    {
      const int nt = tile + gridDim.x;
      if (nt < n_tiles) {
        stage_x1(s.u.x.X1, d.x1 + (size_t)nt * kTE * kRow, min(kTE, d.n_dst - nt * kTE));
        cp_async_commit();
      }
    }
    // ---- F: out = x_dst + D2 + b2 -------------------------------------------------------------------
    {
      const int q = warp & 3, hh = warp >> 2;
      const int row = 32 * q + lane;
      const int node = n0 + (row >> 4);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c0 = 32 * hh + 16 * i;
        float v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + kH + c0, v);
        if (node < d.n_dst) {
          const size_t off = (size_t)node * kRow + (row & 15) * kC + c0;
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const float4 xd = ldg4(d.x_dst + off + e);
            const float4 bb = ld4(s.b2 + c0 + e);
            float4 o = make_float4(xd.x + (v[e] + bb.x), xd.y + (v[e + 1] + bb.y), xd.z + (v[e + 2] + bb.z),
                                   xd.w + (v[e + 3] + bb.w));
            if (d.accumulate_out) {
              const float4 old = ld4(d.out + off + e);
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            st4(d.out + off + e, o);
          }
        }
      }
    }
    tc::tc_fence_before();  // TMEM reads of this tile ordered before the next tile's MMAs (after the next barrier)
    parity ^= 1u;
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace grl

extern "C" int grl_fbconv_node_fwd_tc(const GrlConvDesc* d, grl_stream_t stream) {
  GRL_REQUIRE(d && d->n_dst > 0, GRL_EINVAL, "grl_fbconv_node_fwd_tc: bad descriptor");
  GRL_REQUIRE(d->x1 && d->fiber_kernel && d->bias && d->ln_g && d->ln_b && d->w1 && d->b1 && d->w2 && d->b2 && d->x_dst &&
                  d->out, GRL_EINVAL, "grl_fbconv_node_fwd_tc: null pointer");
  static bool attr = false;
  const int smem = (int)sizeof(grl::NodeTcSmem);
  if (!attr) {
    cudaFuncSetAttribute(grl::fbconv_node_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  const int n_tiles = (d->n_dst + grl::kTE - 1) / grl::kTE;
  int grid = grl::sm_count();
  if (grid > n_tiles) grid = n_tiles;
  grl::fbconv_node_fwd_tc_kernel<<<grid, grl::kThreads, smem, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_fbconv_node_fwd_tc");
}
