// Edge side of the message passing, bf16 tensor-core path (forward): edge basis MLP and the spatial half of the
// fibre-bundle convolution.
//   basis[e] = GELU(GELU(F[e] W1^T + b1) W2^T + b2)          F = 14 polynomial invariants of (i1, i2)
//   x1[d]    = sum_{e in in(d)} (basis[e] Wk^T) * x_src[src(e)]
// Reference lines as in grl_basis.cu / grl_conv_edge.cu (hepi.py:76-82,109-123; ponita/conv.py:71-87,116-149).
// The three contractions run on tcgen05 (bf16 operands, fp32 accumulation in TMEM); the edge basis is stored in
// HBM as bf16 [E][16][64] (2 KB per edge instead of 4); invariants, GELU, the message product and the
// deterministic CSR-ordered segmented sum stay fp32.  Small per-CTA footprints (<= 64 KB smem, <= 128 TMEM
// columns) so 2-3 CTAs share an SM and overlap each other's MMA / epilogue / gather phases.
#include "grl_basis_feat.cuh"
#include "grl_common.cuh"
#include "grl_tc.cuh"

namespace grl {

// ---------------------------------------------------------------------------------------------------
// (E1) edge basis forward
// ---------------------------------------------------------------------------------------------------
struct BasisTcSmem {
  __nv_bfloat16 F[kTM * 16];    // [2 chunks][128 rows][8]: 14 invariants, then (1, 1) multiplying the b1 columns of W1b
  __half H1[kTM * 80];          // [10 chunks][128 rows][8]: GELU(pre1) fp16, chunk 8 = (1, 1, 0 ...), chunk 9 = 0
  __nv_bfloat16 W1b[kC * 16];   // [2 chunks][64 rows n][8 f]: f = 14, 15 hold the bf16 (hi, lo) split of b1
  __half W2h[kC * 80];          // [10 chunks][64 rows n][8 k]: chunk 8 = fp16 (hi, lo) split of b2
  uint64_t bar[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 3) edge_basis_fwd_tc_kernel(const GrlBasisDesc d) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BasisTcSmem& s = *reinterpret_cast<BasisTcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, ch = warp >> 2, row = 32 * q + lane;
  if (tid == 0) {
    tc::mbar_init(&s.bar[0], 1);
    tc::mbar_init(&s.bar[1], 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&s.tmem_base, 128);
  // B operand images from the transposed fp32 weights the descriptor carries (w1t[f][n], w2t[k][n]); both biases ride in
  // the contractions as (hi, lo) pairs against ones columns of the A operands
  for (int i = tid; i < kC * 16; i += kThreads) {
    const int n = i >> 4, f = i & 15;
    float w = d.w1t[f * kC + n];
    if (f >= 14) {
      const float b = d.b1[n];
      const float hi = __bfloat162float(__float2bfloat16_rn(b));
      w = f == 14 ? hi : b - hi;
    }
    s.W1b[tc::op_index(kC, n, f)] = __float2bfloat16_rn(w);
  }
  for (int i = tid; i < kC * 80; i += kThreads) {
    const int n = i / 80, k = i % 80;
    float w = 0.f;
    if (k < kC) {
      w = d.w2t[k * kC + n];
    } else if (k < kC + 2) {
      const float b = d.b2[n];
      const float hi = __half2float(__float2half_rn(b));
      w = k == kC ? hi : b - hi;
    }
    s.W2h[tc::op_index(kC, n, k)] = __float2half_rn(w);
  }
  if (tid < kTM) {  // ones / zero columns of H1 (persistent)
    *reinterpret_cast<uint4*>(s.H1 + ((size_t)8 * kTM + tid) * 8) = make_uint4(0x3C003C00u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(s.H1 + ((size_t)9 * kTM + tid) * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = s.tmem_base, lane_addr = tmem + ((uint32_t)(32 * q) << 16);
  const uint32_t fa = tc::smem_u32(s.F), ha = tc::smem_u32(s.H1), w1 = tc::smem_u32(s.W1b), w2 = tc::smem_u32(s.W2h);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.basis_bf16);
  uint32_t parity = 0;
  const int n_tiles = (d.n_edges + kTE - 1) / kTE;
  BasisFeatPipe feat;
  feat.start(d, blockIdx.x, gridDim.x);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    feat.emit_and_advance(d, tile, gridDim.x, s.F, 1.0f);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      tc::issue_mma(tmem, tc::view_k(fa, kTM), tc::view_k(w1, kC), tc::idesc_bf16(128, kC), 1, false);
      tc::mma_commit(&s.bar[0]);
    }
    tc::mbar_wait(&s.bar[0], parity);
    tc::tc_fence_after();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c0 = 32 * ch + 16 * i;
      float v[16];
      tc::tmem_ld16(lane_addr + c0, v);
      __half2 h0[4], h1[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        h0[e] = tc::gelu_h2(__floats2half2_rn(v[2 * e], v[2 * e + 1]));
        h1[e] = tc::gelu_h2(__floats2half2_rn(v[8 + 2 * e], v[8 + 2 * e + 1]));
      }
      *reinterpret_cast<uint4*>(s.H1 + ((size_t)(c0 >> 3) * kTM + row) * 8) = tc::pack_h8(h0);
      *reinterpret_cast<uint4*>(s.H1 + ((size_t)((c0 >> 3) + 1) * kTM + row) * 8) = tc::pack_h8(h1);
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      tc::issue_mma(tmem + kC, tc::view_k(ha, kTM), tc::view_k(w2, kC), tc::idesc_f16_ex(128, kC, 0, 0, 0, 0), 80 / 16, false);
      tc::mma_commit(&s.bar[1]);
    }
    tc::mbar_wait(&s.bar[1], parity);
    tc::tc_fence_after();
    const int e_idx = tile * kTE + (row >> 4);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c0 = 32 * ch + 16 * i;
      float v[16];
      tc::tmem_ld16(lane_addr + kC + c0, v);
      uint32_t ob[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float2 g = __half22float2(tc::gelu_h2(__floats2half2_rn(v[2 * e], v[2 * e + 1])));
        const __nv_bfloat162 b = __floats2bfloat162_rn(g.x, g.y);
        ob[e] = *reinterpret_cast<const uint32_t*>(&b);
      }
      if (e_idx < d.n_edges) {
        __nv_bfloat16* p = out + (size_t)e_idx * kRow + (row & 15) * kC + c0;
        *reinterpret_cast<uint4*>(p) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
        *reinterpret_cast<uint4*>(p + 8) = make_uint4(ob[4], ob[5], ob[6], ob[7]);
      }
    }
    tc::tc_fence_before();
    parity ^= 1u;
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------------------------------------------
// (E2) spatial convolution forward: kern = basis Wk^T on the tensor core, message product + CSR-ordered sum.
// Persistent CTAs own CONTIGUOUS dst-node ranges of equal cost (edges and nodes; binary search in rowptr), so
// the edge tiles of a CTA are consecutive 8-edge groups of the CSR: tile t+1's basis rows and gathered x_src rows
// are fetched with cp.async into the other half of a 2-stage ring while tile t is multiplied and reduced, and the
// edge indices are prefetched two tiles ahead in registers.
// ---------------------------------------------------------------------------------------------------
struct EdgeFwdTcSmem {
  __nv_bfloat16 BZ[2][kTM * kC];  // basis tile images [8 chunks][128 rows][8]
  float XS[2][kTileFloats];       // gathered x_src rows, then messages in place
  __nv_bfloat16 Wkb[kC * kC];     // [8 chunks][64 rows c][8 j]
  int src[4][kTE], dst[4][kTE];   // 4-deep index ring: slot (t & 3) holds the edges of tile t
  int brow[4][kTE];               // row of each edge in the basis tensor (= edge position unless d.basis_row is given)
  uint64_t bar_x[2];              // transaction barriers of the bulk x_src row gathers, one per ring stage
  uint64_t bar;
  uint32_t tmem_base;
};

// first node n in [0, n_nodes] with cost(n) = 3 * rowptr[n] + 2 * n >= target (an edge moves ~6 KB, a node's x1 row
// 4 KB: edge-less padded nodes still have their zero row written, so ranges are balanced on both)
__device__ __forceinline__ int node_lower_bound(const int32_t* __restrict__ rowptr, int n_nodes, long long target) {
  int lo = 0, hi = n_nodes;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (3ll * __ldg(rowptr + mid) + 2ll * mid < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// basis rows of `cnt` edges (bf16 [16][64] per edge in HBM, edge j at row brow[j]) -> operand image, 16-byte cp.async pieces
__device__ __forceinline__ void stage_basis_image(__nv_bfloat16* __restrict__ img, const __nv_bfloat16* __restrict__ basis,
                                                  const int* __restrict__ brow, int cnt) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int f = threadIdx.x + kThreads * i;  // 16-byte piece 0..1023
    const int r = f >> 3, c8 = f & 7, j = r >> 4;
    __nv_bfloat16* dpt = img + ((size_t)c8 * kTM + r) * 8;
    if (j < cnt) {
      const unsigned sa = tc::smem_u32(dpt);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa),
                   "l"(basis + (size_t)brow[j] * kRow + (r & 15) * kC + 8 * c8) : "memory");
    } else {
      *reinterpret_cast<uint4*>(dpt) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

__global__ void __launch_bounds__(kThreads, 2) fbconv_edge_fwd_tc_kernel(const GrlConvDesc d) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  EdgeFwdTcSmem& s = *reinterpret_cast<EdgeFwdTcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, ch = warp >> 2, row = 32 * q + lane;
  const int o = tid >> 4, cg = tid & 15;  // mapping of the segmented-sum phase
  if (tid == 0) {
    tc::mbar_init(&s.bar, 1);
    tc::mbar_init(&s.bar_x[0], 1);
    tc::mbar_init(&s.bar_x[1], 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&s.tmem_base, 64);
  tc::stage_weight_bf16(s.Wkb, d.wk, kC, kC, kC);
  // this CTA's node range: equal cost shares
  const long long W = 3ll * d.n_edges + 2ll * d.n_dst;
  const int n_lo = blockIdx.x == 0 ? 0 : node_lower_bound(d.rowptr_dst, d.n_dst, W * blockIdx.x / gridDim.x);
  const int n_hi = blockIdx.x + 1 == gridDim.x ? d.n_dst : node_lower_bound(d.rowptr_dst, d.n_dst, W * (blockIdx.x + 1) / gridDim.x);
  const int p0 = d.rowptr_dst[n_lo], p1 = d.rowptr_dst[n_hi];
  const int n_tiles = (p1 - p0 + kTE - 1) / kTE;
  const __nv_bfloat16* basis = reinterpret_cast<const __nv_bfloat16*>(d.basis_bf16);

  // edge indices of tile t live in registers of threads 0..7 until they are published to s.src / s.dst
  auto load_idx = [&](int t, int& es, int& ed, int& eb) {
    es = 0; ed = 0; eb = 0;
    const int e = p0 + t * kTE + tid;
    if (tid < kTE && t < n_tiles && e < p1) {
      es = __ldg(d.edge_src + e);
      ed = __ldg(d.edge_dst + e);
      eb = d.basis_row ? __ldg(d.basis_row + e) : e;  // sub layers read the parent's basis rows in place
    }
  };
  auto stage = [&](int t, int buf) {  // requires index slot (t & 3) to be visible
    const int cnt = min(kTE, p1 - (p0 + t * kTE));
    stage_basis_image(s.BZ[buf], basis, s.brow[t & 3], cnt);
    // x_src rows through the bulk-copy engine: thread r < 128 requests the 256-byte orientation row r of the tile
    // (one cp.async.bulk instead of eight 16-byte cp.async per thread); completion is counted in bytes on bar_x[buf]
    if (tid < kTM) {
      const int j = tid >> 4, oo = tid & 15;
      float* drow = s.XS[buf] + tid * kLDT;
      if (j < cnt) {
        tc::bulk_g2s(drow, d.x_src + (size_t)s.src[t & 3][j] * kRow + oo * kC, kC * 4u, &s.bar_x[buf]);
      } else {
#pragma unroll
        for (int c = 0; c < kC; c += 4) *reinterpret_cast<float4*>(drow + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (tid == 0) tc::mbar_expect_tx(&s.bar_x[buf], (uint32_t)cnt * kO * kC * 4u);
    }
  };
  int es_a, ed_a, eb_a, es_b, ed_b, eb_b;  // a: tile t+2 (published in iteration t), b: tile t+3
  load_idx(0, es_a, ed_a, eb_a);
  if (tid < kTE) { s.src[0][tid] = es_a; s.dst[0][tid] = ed_a; s.brow[0][tid] = eb_a; }
  load_idx(1, es_a, ed_a, eb_a);
  if (tid < kTE) { s.src[1][tid] = es_a; s.dst[1][tid] = ed_a; s.brow[1][tid] = eb_a; }
  load_idx(2, es_a, ed_a, eb_a);
  load_idx(3, es_b, ed_b, eb_b);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = s.tmem_base, lane_addr = tmem + ((uint32_t)(32 * q) << 16);
  const uint32_t wk = tc::smem_u32(s.Wkb);
  if (n_tiles > 0) stage(0, 0);
  cp_async_commit();

  uint32_t parity = 0;
  int cur = n_lo;
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    const int cnt = min(kTE, p1 - (p0 + t * kTE));
    // publish the indices of tile t+2: slot (t+2)&3 was last read by tile t-2, two barriers ago
    if (tid < kTE) { s.src[(t + 2) & 3][tid] = es_a; s.dst[(t + 2) & 3][tid] = ed_a; s.brow[(t + 2) & 3][tid] = eb_a; }
    es_a = es_b; ed_a = ed_b; eb_a = eb_b;
    load_idx(t + 4, es_b, ed_b, eb_b);
    tc::fence_async_smem();  // generic-proxy accesses to XS[buf ^ 1] (tile t-1) before the bulk engine rewrites it
    __syncthreads();  // everyone is done reducing tile t-1: its data buffers (buf ^ 1) may be overwritten
    if (t + 1 < n_tiles) stage(t + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();  // tile t has landed (this thread's copies); the barrier below makes it CTA-wide
    tc::mbar_wait(&s.bar_x[buf], (uint32_t)(t >> 1) & 1u);  // ... and so have its x_src rows
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      tc::issue_mma(tmem, tc::view_k(tc::smem_u32(s.BZ[buf]), kTM), tc::view_k(wk, kC), tc::idesc_bf16(128, kC), kC / 16, false);
      tc::mma_commit(&s.bar);
    }
    tc::mbar_wait(&s.bar, parity);
    parity ^= 1u;
    tc::tc_fence_after();
#pragma unroll
    for (int i = 0; i < 2; ++i) {  // messages in place: XS[row][c] *= kern[row][c]
      const int c0 = 32 * ch + 16 * i;
      float v[16];
      tc::tmem_ld16(lane_addr + c0, v);
      float* xs = s.XS[buf] + row * kLDT + c0;
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        float4 x = ld4(xs + e);
        x.x *= v[e]; x.y *= v[e + 1]; x.z *= v[e + 2]; x.w *= v[e + 3];
        st4(xs + e, x);
      }
    }
    tc::tc_fence_before();
    __syncthreads();
    // CSR-ordered segmented sum: thread (o, 4 channels) adds the edges of the tile sequentially
#pragma unroll
    for (int j = 0; j < kTE; ++j) {
      if (j < cnt) {
        const int dn = s.dst[t & 3][j];
        while (cur < dn) {
          st4(d.x1 + (size_t)cur * kRow + o * kC + 4 * cg, sum);
          sum = make_float4(0.f, 0.f, 0.f, 0.f);
          ++cur;
        }
        const float4 m = ld4(s.XS[buf] + (16 * j + o) * kLDT + 4 * cg);
        sum.x += m.x; sum.y += m.y; sum.z += m.z; sum.w += m.w;
      }
    }
  }
  while (cur < n_hi) {
    st4(d.x1 + (size_t)cur * kRow + o * kC + 4 * cg, sum);
    sum = make_float4(0.f, 0.f, 0.f, 0.f);
    ++cur;
  }
  cp_async_wait_all();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

}  // namespace grl

extern "C" {

int grl_edge_basis_fwd_tc(const GrlBasisDesc* d, grl_stream_t stream) {
  GRL_REQUIRE(d, GRL_EINVAL, "grl_edge_basis_fwd_tc: null descriptor");
  if (d->n_edges == 0) return GRL_OK;
  GRL_REQUIRE(d->n_edges > 0 && (d->dim == 2 || d->dim == 3), GRL_EINVAL, "grl_edge_basis_fwd_tc: n_edges=%d dim=%d",
              d->n_edges, d->dim);
  GRL_REQUIRE(d->edge_src && d->edge_dst && d->pos_src && d->pos_dst && d->ori && d->w1t && d->b1 && d->w2t && d->b2 &&
                  d->basis_bf16, GRL_EINVAL, "grl_edge_basis_fwd_tc: null pointer");
  const int smem = (int)sizeof(grl::BasisTcSmem);
  if (grl::ensure_dynamic_smem((const void*)grl::edge_basis_fwd_tc_kernel, smem) != GRL_OK) return GRL_ECUDA;
  const int n_tiles = (d->n_edges + grl::kTE - 1) / grl::kTE;
  int grid = 3 * grl::sm_count();  // 37 KB smem, 128 TMEM columns, <= 85 registers: three CTAs share an SM
  if (grid > n_tiles) grid = n_tiles;
  grl::edge_basis_fwd_tc_kernel<<<grid, grl::kThreads, smem, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_edge_basis_fwd_tc");
}

int grl_fbconv_edge_fwd_tc(const GrlConvDesc* d, grl_stream_t stream) {
  GRL_REQUIRE(d, GRL_EINVAL, "grl_fbconv_edge_fwd_tc: null descriptor");
  GRL_REQUIRE(d->n_dst > 0 && d->n_src > 0 && d->n_edges >= 0, GRL_EINVAL, "grl_fbconv_edge_fwd_tc: bad sizes");
  GRL_REQUIRE(d->rowptr_dst && d->x_src && d->wk && d->x1 && (d->n_edges == 0 || (d->edge_src && d->edge_dst && d->basis_bf16)),
              GRL_EINVAL, "grl_fbconv_edge_fwd_tc: null pointer");
  const int smem = (int)sizeof(grl::EdgeFwdTcSmem);
  if (grl::ensure_dynamic_smem((const void*)grl::fbconv_edge_fwd_tc_kernel, smem) != GRL_OK) return GRL_ECUDA;
  int grid = 2 * grl::sm_count();
  if (grid > d->n_dst) grid = d->n_dst;
  grl::fbconv_edge_fwd_tc_kernel<<<grid, grl::kThreads, smem, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_fbconv_edge_fwd_tc");
}

}  // extern "C"
