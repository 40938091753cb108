// Edge side of the message passing, bf16 tensor-core path (backward).
//
// (E3) fbconv_edge_bwd_tc_kernel — src-sorted tiles of 8 edges:
//        kern    = basis Wk^T                         (tcgen05)
//        g_xsrc[s] = init[s] + sum_{e in out(s)} g_x1[dst(e)] * kern[e]       (fp32, src-CSR order, no atomics)
//        g_kern  = g_x1[dst(e)] * x_src[src(e)]
//        g_basis = g_kern Wk                          (tcgen05, Wk image read MN-major)  -> bf16 in HBM
//        gWk    += g_kern^T basis                     (tcgen05, accumulated in TMEM over all tiles of the CTA)
// (E4) edge_basis_bwd_tc_kernel — edge-order tiles: recompute F, pre1, H1, pre2, then
//        gP2 = g_basis * GELU'(pre2);  gW2 += gP2^T H1;  gH1 = gP2 W2;  gP1 = gH1 * GELU'(pre1);  gW1 += gP1^T F
//
// The 64 x 64 (and 64 x 16) weight gradients are formed with M = 128 MMAs by reading a 128-column operand image
// made of two ADJACENT 64-column images [X | G]: lanes 64..127 of the accumulator then hold G^T Y (lanes 0..63
// hold X^T Y and are ignored).  Reference: autograd of ponita/conv.py:84-87,116-149 and hepi.py:76-82,109-123.
#include "grl_basis_feat.cuh"
#include "grl_common.cuh"
#include "grl_tc.cuh"

namespace grl {

__device__ __forceinline__ void cp_async16_any(void* smem_dst, const void* gmem_src) {
  const unsigned sa = tc::smem_u32(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem_src) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// (E3) structured around latency:
//   * persistent CTAs own CONTIGUOUS src-node ranges holding equal shares of the edges (binary search in
//     rowptr_src), so a CTA walks consecutive 8-entry tiles of the src-sorted list;
//   * the (eid, src, dst) triples need two dependent global loads: they are fetched two tiles ahead into
//     registers of threads 0..7 and published through a 4-deep shared-memory ring;
//   * one tile ahead, threads 0..31 pull the rows the next tile will gather (basis 2 KB, g_x1 4 KB, x_src 4 KB per
//     edge) into L2 with cp.async.bulk.prefetch.L2, so the cp.async gathers pay L2 hits, not DRAM round trips;
//   * the residual row added at a node's flush (grad_x_src_init) is loaded when the node STARTS accumulating.
// ncu r01 of the block-interleaved, unpipelined version: smsp__issue_active 19 %, stall samples on the index loads (8 %), the gather issue and
// wait (36 %) and the flush's dependent global load (14 %).
// ---------------------------------------------------------------------------------------------------
struct EdgeBwd2Smem {
  __nv_bfloat16 BZ[kTM * kC];   // basis rows gathered by edge id   [8 chunks][128][8]
  __nv_bfloat16 GK[kTM * kC];   // g_kern, MUST directly follow BZ ([BZ | GK] is read as one 128-column image)
  __nv_bfloat16 Wkb[kC * kC];   // [8 chunks][64 rows c][8 j]
  float GX[kTileFloats];        // g_x1 rows gathered by dst, then g_x1 * kern in place
  float XS[kTileFloats];        // x_src rows gathered by src
  int eid[4][kTE], src[4][kTE], dst[4][kTE];  // index ring: slot (t & 3) holds tile t
  int acc[4][kTE];              // per edge: add to the basis-gradient row already in HBM (1) or overwrite it (0)
  int lead[4][kTE];             // first edge of the tile with the same source node: its x_src row is staged once and shared
  uint64_t bar[2];
  uint64_t bar_g;               // transaction barrier of the bulk row gathers (GX, XS)
  uint32_t tmem_base;
};

// The fp32 row gathers (g_x1 by dst, x_src by src) go through the bulk-copy engine: one `cp.async.bulk` per 256-byte
// orientation row instead of sixteen 16-byte cp.async per thread, which was 35 % of the kernel's instructions and 40 %
// of its stall samples (profiles/r01b_full.md).  0 = the cp.async path.
#ifndef GRL_BULK_GATHER
#define GRL_BULK_GATHER 1
#endif

// first node n in [0, n_nodes] with cost(n) = 2 * rowptr[n] + n >= target: a node costs about half an edge (its
// flush moves 8 KB, an edge ~12 KB plus the MMAs), so ranges are balanced on edges AND nodes (padded / isolated
// nodes have no edges but still need their residual row copied)
__device__ __forceinline__ int node_lower_bound_src(const int32_t* __restrict__ rowptr, int n_nodes, long long target) {
  int lo = 0, hi = n_nodes;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (2ll * __ldg(rowptr + mid) + mid < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// stage_rows_gather for a src-sorted tile: only the first edge of a run of equal sources copies its row (the others read
// the leader's copy); rows past `cnt` are zero-filled (they meet zero g_x1 rows in the g_kern product, 0 * garbage != 0)
__device__ __forceinline__ void stage_rows_gather_lead(float* __restrict__ tile, const float* __restrict__ base,
                                                       const int* __restrict__ idx, const int* __restrict__ lead, int cnt) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = threadIdx.x + kThreads * i;
    const int row = f >> 4, c4 = f & 15;
    const int j = row >> 4;
    float* d = tile + row * kLDT + 4 * c4;
    if (j < cnt) {
      if (lead[j] == j) cp_async16(d, base + (size_t)idx[j] * kRow + (row & 15) * kC + 4 * c4);
    } else {
      *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

__global__ void __launch_bounds__(kThreads, 2) fbconv_edge_bwd_tc2_kernel(const GrlConvDesc d) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  EdgeBwd2Smem& s = *reinterpret_cast<EdgeBwd2Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, ch = warp >> 2, row = 32 * q + lane;
  const int o = tid >> 4, cg = tid & 15;
  if (tid == 0) {
    tc::mbar_init(&s.bar[0], 1);
    tc::mbar_init(&s.bar[1], 1);
    tc::mbar_init(&s.bar_g, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&s.tmem_base, 256);
  tc::stage_weight_bf16(s.Wkb, d.wk, kC, kC, kC);
  // this CTA's src-node range: equal cost shares
  const long long W = 2ll * d.n_edges + d.n_src;  // total cost
  const int n_lo = blockIdx.x == 0 ? 0 : node_lower_bound_src(d.rowptr_src, d.n_src, W * blockIdx.x / gridDim.x);
  const int n_hi = blockIdx.x + 1 == gridDim.x ? d.n_src : node_lower_bound_src(d.rowptr_src, d.n_src, W * (blockIdx.x + 1) / gridDim.x);
  const int p0 = d.rowptr_src[n_lo], p1 = d.rowptr_src[n_hi];
  const int n_tiles = (p1 - p0 + kTE - 1) / kTE;
  const __nv_bfloat16* basis = reinterpret_cast<const __nv_bfloat16*>(d.basis_bf16);
  __nv_bfloat16* g_basis = reinterpret_cast<__nv_bfloat16*>(d.grad_basis_bf16);

  // Index pipeline of threads 0..7 (no load ever waits on a load issued in the same iteration):
  //   iteration t issues  eid(t+4) = src_eid[..]            and  (src, dst)(t+3) = edge_{src,dst}[eid(t+3)],
  //   and publishes tile t+2, whose three values were requested in iterations t-2 and t-1.
  auto load_eid = [&](int t) -> int {
    const int p = p0 + t * kTE + tid;
    return (tid < kTE && t < n_tiles && p < p1) ? __ldg(d.src_eid + p) : -1;
  };
  // accumulate_grad_basis: 0 = overwrite every row, 1 = add to every row, 2 = add where grad_basis_acc_mask[e] >= 0
  // (rows a sub layer of the same basis wrote before this launch), overwrite elsewhere
  const int acc_mode = d.accumulate_grad_basis;
  auto load_sd = [&](int e, int& es, int& ed, int& ea) {
    es = 0; ed = 0; ea = acc_mode == 1;
    if (e >= 0) {
      es = __ldg(d.edge_src + e);
      ed = __ldg(d.edge_dst + e);
      if (acc_mode == 2) ea = __ldg(d.grad_basis_acc_mask + e) >= 0;
    }
  };
  auto publish = [&](int slot, int e, int es, int ed, int ea) {
    if (warp == 0) {  // src-sorted list: equal sources are adjacent, so a run inside the tile shares one staged x_src row
      const int prev = __shfl_up_sync(0xffffffffu, es, 1);
      const bool starts = lane == 0 || lane >= kTE || es != prev;
      const unsigned heads = __ballot_sync(0xffffffffu, starts);
      const int ld = 31 - __clz(heads & ((2u << lane) - 1u));
      if (tid < kTE) {
        s.eid[slot][tid] = e < 0 ? 0 : e; s.src[slot][tid] = es; s.dst[slot][tid] = ed; s.acc[slot][tid] = ea;
        s.lead[slot][tid] = ld;
      }
    }
  };
  int e_2, es_2, ed_2, ea_2;  // tile t+2: complete
  int e_3;                    // tile t+3: eid requested, (src, dst) not yet
  {
    int e0 = load_eid(0), e1 = load_eid(1), es, ed, ea;
    e_2 = load_eid(2);
    e_3 = load_eid(3);
    load_sd(e0, es, ed, ea);
    publish(0, e0, es, ed, ea);
    load_sd(e1, es, ed, ea);
    publish(1, e1, es, ed, ea);
    load_sd(e_2, es_2, ed_2, ea_2);
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = s.tmem_base, lane_addr = tmem + ((uint32_t)(32 * q) << 16);
  const uint32_t bz = tc::smem_u32(s.BZ), gk = tc::smem_u32(s.GK), wk = tc::smem_u32(s.Wkb);
  uint32_t par0 = 0, par1 = 0, par_g = 0;
  bool first = true;

  // rows of tile t -> L2 (threads 0..23: 8 edges x {basis, g_x1, x_src}); slot (t & 3) must be visible
  auto prefetch_tile = [&](int t) {
    if (tid < 4 * kTE && t < n_tiles) {
      const int j = tid & 7, which = tid >> 3;
      if (p0 + t * kTE + j < p1) {
        if (which == 0) tc::prefetch_l2(basis + (size_t)s.eid[t & 3][j] * kRow, kRow * 2u);
        else if (which == 1) tc::prefetch_l2(d.grad_x1 + (size_t)s.dst[t & 3][j] * kRow, kRow * 4u);
        else if (which == 2) { if (s.lead[t & 3][j] == j) tc::prefetch_l2(d.x_src + (size_t)s.src[t & 3][j] * kRow, kRow * 4u); }
        else if (s.acc[t & 3][j]) tc::prefetch_l2(g_basis + (size_t)s.eid[t & 3][j] * kRow, kRow * 2u);
      }
    }
  };
  prefetch_tile(0);

  int cur = n_lo;
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 init = make_float4(0.f, 0.f, 0.f, 0.f);
  auto begin_node = [&](int node) {  // residual row of `node`, added at its flush
    if (d.grad_x_src_init && node < n_hi) init = ldg4(d.grad_x_src_init + (size_t)node * kRow + o * kC + 4 * cg);
  };
  auto flush = [&](int node) {
    sum.x += init.x; sum.y += init.y; sum.z += init.z; sum.w += init.w;
    st4(d.grad_x_src + (size_t)node * kRow + o * kC + 4 * cg, sum);
    sum = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  // advance to node `to`: flush the current node, then copy the residual rows of the edge-less nodes in between
  // four at a time (their loads are independent, so a run of padded nodes costs one latency per four nodes)
  auto advance = [&](int to) {
    flush(cur);
    ++cur;
    const size_t toff = (size_t)o * kC + 4 * cg;
    while (cur < to) {
      const int n = min(4, to - cur);
      float4 r[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        r[i] = (d.grad_x_src_init && i < n) ? ldg4(d.grad_x_src_init + (size_t)(cur + i) * kRow + toff) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < n) st4(d.grad_x_src + (size_t)(cur + i) * kRow + toff, r[i]);
      cur += n;
    }
    begin_node(cur);
  };
  begin_node(cur);

  for (int t = 0; t < n_tiles; ++t) {
    const int slot = t & 3;
    const int base = p0 + t * kTE;
    const int cnt = min(kTE, p1 - base);
    publish((t + 2) & 3, e_2, es_2, ed_2, ea_2);  // slot (t+2)&3 was last read by tile t-2, two barriers ago
    e_2 = e_3;
    load_sd(e_2, es_2, ed_2, ea_2);               // eid(t+3) was requested one iteration ago
    e_3 = load_eid(t + 4);
#if GRL_BULK_GATHER
    tc::fence_async_smem();  // this thread's generic-proxy accesses to GX / XS (tile t-1) before the async-proxy rewrites
#endif
    __syncthreads();  // everyone is done with tile t-1's buffers; slot (t+1)&3 (published last iteration) is visible
    prefetch_tile(t + 1);
    if (tid == 32 && d.grad_x_src_init && cur + 2 < n_hi)  // residual rows of the next few nodes -> L2
      tc::prefetch_l2(d.grad_x_src_init + (size_t)(cur + 2) * kRow, (uint32_t)min(8, n_hi - cur - 2) * kRow * 4u);
    // gathers of tile t: basis rows (bf16) by edge id -> operand image, g_x1 rows by dst, x_src rows by src
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int f = tid + kThreads * i;
      const int r = f >> 3, c8 = f & 7, j = r >> 4;
      __nv_bfloat16* dpt = s.BZ + ((size_t)c8 * kTM + r) * 8;
      if (j < cnt) cp_async16_any(dpt, basis + (size_t)s.eid[slot][j] * kRow + (r & 15) * kC + 8 * c8);
      else *reinterpret_cast<uint4*>(dpt) = make_uint4(0u, 0u, 0u, 0u);
    }
#if GRL_BULK_GATHER
    {
      // thread r < 128 owns row r of GX (g_x1 by dst), thread 128 + r row r of XS (x_src by src, run leaders only)
      const int rr = tid & 127, j = rr >> 4, oo = rr & 15;
      const bool is_x = tid >= 128;
      float* drow = (is_x ? s.XS : s.GX) + rr * kLDT;
      if (j < cnt) {
        if (!is_x) tc::bulk_g2s(drow, d.grad_x1 + (size_t)s.dst[slot][j] * kRow + oo * kC, kC * 4u, &s.bar_g);
        else if (s.lead[slot][j] == j) tc::bulk_g2s(drow, d.x_src + (size_t)s.src[slot][j] * kRow + oo * kC, kC * 4u, &s.bar_g);
      } else {  // rows past the end of the list: zeros (they meet zero rows in the products, 0 * garbage != 0)
#pragma unroll
        for (int c = 0; c < kC; c += 4) *reinterpret_cast<float4*>(drow + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (tid == 0) {
        int n_lead = 0;
        for (int jj = 0; jj < cnt; ++jj) n_lead += s.lead[slot][jj] == jj;
        tc::mbar_expect_tx(&s.bar_g, (uint32_t)(cnt + n_lead) * kO * kC * 4u);
      }
    }
    cp_async_commit();
    cp_async_wait_all();
    tc::mbar_wait(&s.bar_g, par_g);
    par_g ^= 1u;
#else
    stage_rows_gather(s.GX, d.grad_x1, s.dst[slot], cnt);
    stage_rows_gather_lead(s.XS, d.x_src, s.src[slot], s.lead[slot], cnt);
    cp_async_commit();
    cp_async_wait_all();
#endif
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      tc::issue_mma(tmem, tc::view_k(bz, kTM), tc::view_k(wk, kC), tc::idesc_bf16(128, kC), kC / 16, false);
      tc::mma_commit(&s.bar[0]);
    }
    tc::mbar_wait(&s.bar[0], par0);
    par0 ^= 1u;
    tc::tc_fence_after();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c0 = 32 * ch + 16 * i;
      float v[16], gkv[16];
      tc::tmem_ld16(lane_addr + c0, v);
      float* gx = s.GX + row * kLDT + c0;
      const float* xs = s.XS + (16 * s.lead[slot][row >> 4] + (row & 15)) * kLDT + c0;
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        const float4 gm = ld4(gx + e), x = ld4(xs + e);
        st4(gx + e, make_float4(gm.x * v[e], gm.y * v[e + 1], gm.z * v[e + 2], gm.w * v[e + 3]));
        gkv[e] = gm.x * x.x; gkv[e + 1] = gm.y * x.y; gkv[e + 2] = gm.z * x.z; gkv[e + 3] = gm.w * x.w;
      }
      *reinterpret_cast<uint4*>(s.GK + ((size_t)(c0 >> 3) * kTM + row) * 8) = tc::pack8(gkv);
      *reinterpret_cast<uint4*>(s.GK + ((size_t)((c0 >> 3) + 1) * kTM + row) * 8) = tc::pack8(gkv + 8);
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      // g_basis = g_kern Wk   (Wk image [c rows][j cols] read MN-major: K = c, N = j)
      tc::issue_mma(tmem + 64, tc::view_k(gk, kTM), tc::view_mn(wk, kC), tc::idesc_bf16_ex(128, 64, 0, 1), kC / 16, false);
      // lanes 64..127: gWk[c][j] += sum_rows g_kern[row][c] basis[row][j]   ([BZ | GK] as one 128-column image)
      tc::issue_mma(tmem + 128, tc::view_mn(bz, kTM), tc::view_mn(bz, kTM), tc::idesc_bf16_ex(128, 64, 1, 1), kTM / 16, !first);
      tc::mma_commit(&s.bar[1]);
    }
    first = false;
    // another layer's share of the basis gradient (accumulate mode): requested before the segmented sum, used after the MMA wait
    uint4 og[4];
    const bool acc_row = (row >> 4) < cnt && s.acc[slot][row >> 4] != 0;
    if (acc_row) {
      const uint4* p = reinterpret_cast<const uint4*>(g_basis + (size_t)s.eid[slot][row >> 4] * kRow + (row & 15) * kC + 32 * ch);
#pragma unroll
      for (int i = 0; i < 4; ++i) og[i] = p[i];
    }
    // src-CSR segmented sum of g_x1 * kern while the tensor core works
#pragma unroll
    for (int j = 0; j < kTE; ++j) {
      if (j < cnt) {
        const int sn = s.src[slot][j];
        if (cur < sn) advance(sn);
        const float4 m = ld4(s.GX + (16 * j + o) * kLDT + 4 * cg);
        sum.x += m.x; sum.y += m.y; sum.z += m.z; sum.w += m.w;
      }
    }
    tc::mbar_wait(&s.bar[1], par1);
    par1 ^= 1u;
    tc::tc_fence_after();
    {
      const int j = row >> 4;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c0 = 32 * ch + 16 * i;
        float v[16];
        tc::tmem_ld16(lane_addr + 64 + c0, v);
        if (j < cnt) {
          __nv_bfloat16* p = g_basis + (size_t)s.eid[slot][j] * kRow + (row & 15) * kC + c0;
          if (acc_row) {  // add in fp32, round once
            const uint4 o0 = og[2 * i], o1 = og[2 * i + 1];
            const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&o0);
            const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&o1);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 a = __bfloat1622float2(h0[e]), b = __bfloat1622float2(h1[e]);
              v[2 * e] += a.x; v[2 * e + 1] += a.y;
              v[8 + 2 * e] += b.x; v[8 + 2 * e + 1] += b.y;
            }
          }
          *reinterpret_cast<uint4*>(p) = tc::pack8(v);
          *reinterpret_cast<uint4*>(p + 8) = tc::pack8(v + 8);
        }
      }
    }
    tc::tc_fence_before();
  }
  if (cur < n_hi) advance(n_hi);
  // gWk partial of this CTA: lanes 64..127 = rows c, columns j
  __syncthreads();
  tc::tc_fence_after();
  float* P = d.edge_grad_partials + (size_t)blockIdx.x * kWFloats;
#pragma unroll 1
  for (int i = 0; i < 2; ++i) {
    const int c0 = 32 * ch + 16 * i;
    float v[16];
    if (first) {
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] = 0.f;
    } else {
      tc::tmem_ld16(lane_addr + 128 + c0, v);
    }
    if (row >= 64) {
      float* p = P + (size_t)(row - 64) * kC + c0;
#pragma unroll
      for (int e = 0; e < 16; e += 4) st4(p + e, make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]));
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------------
// (E4)
// ---------------------------------------------------------------------------------------------------
struct BasisBwdTcSmem {
  __nv_bfloat16 F[kTM * 16];    // [2 chunks][128][8]
  __nv_bfloat16 H1[kTM * kC];   // [8 chunks][128][8]
  __nv_bfloat16 GP2[kTM * kC];  // MUST follow H1   ([H1 | GP2] -> gW2)
  __nv_bfloat16 GP1[kTM * kC];  // MUST follow GP2  ([GP2 | GP1] -> gW1)
  __nv_bfloat16 W1b[kC * 16];
  __nv_bfloat16 W2b[kC * kC];
  float b1[kC], b2[kC];
  float acc_gb1[4][kC], acc_gb2[4][kC];
  uint64_t bar[4];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 2) edge_basis_bwd_tc_kernel(const GrlBasisDesc d) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BasisBwdTcSmem& s = *reinterpret_cast<BasisBwdTcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, ch = warp >> 2, row = 32 * q + lane;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) tc::mbar_init(&s.bar[i], 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&s.tmem_base, 256);
  for (int i = tid; i < kC * 16; i += kThreads) {
    const int n = i >> 4, f = i & 15;
    s.W1b[tc::op_index(kC, n, f)] = __float2bfloat16(d.w1t[f * kC + n]);
  }
  for (int i = tid; i < kC * kC; i += kThreads) {
    const int n = i >> 6, k = i & 63;
    s.W2b[tc::op_index(kC, n, k)] = __float2bfloat16(d.w2t[k * kC + n]);
  }
  if (tid < kC) { s.b1[tid] = d.b1[tid]; s.b2[tid] = d.b2[tid]; }
  for (int i = tid; i < 4 * kC; i += kThreads) { (&s.acc_gb1[0][0])[i] = 0.f; (&s.acc_gb2[0][0])[i] = 0.f; }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = s.tmem_base, lane_addr = tmem + ((uint32_t)(32 * q) << 16);
  const uint32_t fa = tc::smem_u32(s.F), ha = tc::smem_u32(s.H1), g2a = tc::smem_u32(s.GP2), g1a = tc::smem_u32(s.GP1);
  const uint32_t w1 = tc::smem_u32(s.W1b), w2 = tc::smem_u32(s.W2b);
  const __nv_bfloat16* gb = reinterpret_cast<const __nv_bfloat16*>(d.grad_basis_bf16);
  uint32_t parity = 0;
  bool first = true;
  const int n_tiles = (d.n_edges + kTE - 1) / kTE;
  BasisFeatPipe feat;
  feat.start(d, blockIdx.x, gridDim.x);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    feat.emit_and_advance(d, tile, gridDim.x, s.F, 0.0f);
    if (tid == 0 && tile + (int)gridDim.x < n_tiles) {  // next tile's grad_basis rows (contiguous) -> L2
      const int nt = tile + gridDim.x;
      tc::prefetch_l2(gb + (size_t)nt * kTE * kRow, (uint32_t)min(kTE, d.n_edges - nt * kTE) * kRow * 2u);
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {  // pre1 = F W1^T
      tc::tc_fence_after();
      tc::issue_mma(tmem, tc::view_k(fa, kTM), tc::view_k(w1, kC), tc::idesc_bf16(128, kC), 1, false);
      tc::mma_commit(&s.bar[0]);
    }
    tc::mbar_wait(&s.bar[0], parity);
    tc::tc_fence_after();
    float dG1[32];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c0 = 32 * ch + 16 * i;
      float v[16];
      tc::tmem_ld16(lane_addr + c0, v);
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        gelu_fast(v[e] + s.b1[c0 + e], v[e], dG1[16 * i + e]);
      }
      *reinterpret_cast<uint4*>(s.H1 + ((size_t)(c0 >> 3) * kTM + row) * 8) = tc::pack8(v);
      *reinterpret_cast<uint4*>(s.H1 + ((size_t)((c0 >> 3) + 1) * kTM + row) * 8) = tc::pack8(v + 8);
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {  // pre2 = H1 W2^T
      tc::tc_fence_after();
      tc::issue_mma(tmem + 64, tc::view_k(ha, kTM), tc::view_k(w2, kC), tc::idesc_bf16(128, kC), kC / 16, false);
      tc::mma_commit(&s.bar[1]);
    }
    // this thread's grad_basis piece: requested before the wait so the round trip hides behind the pre2 MMA
    // (the tile's rows were pulled into L2 one tile ahead)
    uint4 gq[4];
    {
      const int e_idx = tile * kTE + (row >> 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) gq[i] = make_uint4(0u, 0u, 0u, 0u);
      if (e_idx < d.n_edges) {
        const uint4* p = reinterpret_cast<const uint4*>(gb + (size_t)e_idx * kRow + (row & 15) * kC + 32 * ch);
#pragma unroll
        for (int i = 0; i < 4; ++i) gq[i] = __ldg(p + i);
      }
    }
    tc::mbar_wait(&s.bar[1], parity);
    tc::tc_fence_after();
    {
      float gp2[32];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c0 = 32 * ch + 16 * i;
        float v[16];
        tc::tmem_ld16(lane_addr + 64 + c0, v);
        const uint4 g0 = gq[2 * i], g1 = gq[2 * i + 1];
        const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&g0);
        const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&g1);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = __bfloat1622float2(h0[e]), b = __bfloat1622float2(h1[e]);
          float y_, d0, d1, d2, d3;
          gelu_fast(v[2 * e] + s.b2[c0 + 2 * e], y_, d0);
          gelu_fast(v[2 * e + 1] + s.b2[c0 + 2 * e + 1], y_, d1);
          gelu_fast(v[8 + 2 * e] + s.b2[c0 + 8 + 2 * e], y_, d2);
          gelu_fast(v[8 + 2 * e + 1] + s.b2[c0 + 8 + 2 * e + 1], y_, d3);
          gp2[16 * i + 2 * e] = a.x * d0;
          gp2[16 * i + 2 * e + 1] = a.y * d1;
          gp2[16 * i + 8 + 2 * e] = b.x * d2;
          gp2[16 * i + 8 + 2 * e + 1] = b.y * d3;
        }
        *reinterpret_cast<uint4*>(s.GP2 + ((size_t)(c0 >> 3) * kTM + row) * 8) = tc::pack8(gp2 + 16 * i);
        *reinterpret_cast<uint4*>(s.GP2 + ((size_t)((c0 >> 3) + 1) * kTM + row) * 8) = tc::pack8(gp2 + 16 * i + 8);
      }
      tc::warp_colsum<32>(gp2, lane);
      s.acc_gb2[q][32 * ch + lane] += gp2[0];
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      // gH1 = gP2 W2   (W2 image [n rows][k cols] read MN-major: K = n, N = k)
      tc::issue_mma(tmem, tc::view_k(g2a, kTM), tc::view_mn(w2, kC), tc::idesc_bf16_ex(128, 64, 0, 1), kC / 16, false);
      // lanes 64..127: gW2[n][k] += sum_rows gP2[row][n] H1[row][k]      ([H1 | GP2] as one 128-column image)
      tc::issue_mma(tmem + 128, tc::view_mn(ha, kTM), tc::view_mn(ha, kTM), tc::idesc_bf16_ex(128, 64, 1, 1), kTM / 16, !first);
      tc::mma_commit(&s.bar[2]);
    }
    tc::mbar_wait(&s.bar[2], parity);
    tc::tc_fence_after();
    {
      float gp1[32];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c0 = 32 * ch + 16 * i;
        float v[16];
        tc::tmem_ld16(lane_addr + c0, v);
#pragma unroll
        for (int e = 0; e < 16; ++e) gp1[16 * i + e] = v[e] * dG1[16 * i + e];
        *reinterpret_cast<uint4*>(s.GP1 + ((size_t)(c0 >> 3) * kTM + row) * 8) = tc::pack8(gp1 + 16 * i);
        *reinterpret_cast<uint4*>(s.GP1 + ((size_t)((c0 >> 3) + 1) * kTM + row) * 8) = tc::pack8(gp1 + 16 * i + 8);
      }
      tc::warp_colsum<32>(gp1, lane);
      s.acc_gb1[q][32 * ch + lane] += gp1[0];
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      // lanes 64..127: gW1[n][f] += sum_rows gP1[row][n] F[row][f]       ([GP2 | GP1] as one 128-column image)
      tc::issue_mma(tmem + 192, tc::view_mn(g2a, kTM), tc::view_mn(fa, kTM), tc::idesc_bf16_ex(128, 16, 1, 1), kTM / 16, !first);
      tc::mma_commit(&s.bar[3]);
    }
    tc::mbar_wait(&s.bar[3], parity);
    tc::tc_fence_after();
    tc::tc_fence_before();
    first = false;
    parity ^= 1u;
  }
  // partial slot: gW1[64][16] | gb1[64] | gW2[64][64] | gb2[64]
  __syncthreads();
  tc::tc_fence_after();
  float* P = d.grad_partials + (size_t)blockIdx.x * GRL_BASIS_GRAD_FLOATS;
  {
    float v[16];
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {
      const int c0 = 32 * ch + 16 * i;
      if (first) {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = 0.f;
      } else {
        tc::tmem_ld16(lane_addr + 128 + c0, v);
      }
      if (row >= 64) {
        float* p = P + 64 * 16 + 64 + (size_t)(row - 64) * kC + c0;
#pragma unroll
        for (int e = 0; e < 16; e += 4) st4(p + e, make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]));
      }
    }
    if (!first) tc::tmem_ld16(lane_addr + 192, v);
    if (row >= 64 && ch == 0) {
      float* p = P + (size_t)(row - 64) * 16;
#pragma unroll
      for (int e = 0; e < 16; e += 4) st4(p + e, make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]));
    }
  }
  if (tid < kC) {
    P[64 * 16 + tid] = ((s.acc_gb1[0][tid] + s.acc_gb1[1][tid]) + s.acc_gb1[2][tid]) + s.acc_gb1[3][tid];
    P[64 * 16 + 64 + kWFloats + tid] = ((s.acc_gb2[0][tid] + s.acc_gb2[1][tid]) + s.acc_gb2[2][tid]) + s.acc_gb2[3][tid];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace grl

extern "C" {

int grl_fbconv_edge_bwd_tc(const GrlConvDesc* d, grl_stream_t stream) {
  GRL_REQUIRE(d, GRL_EINVAL, "grl_fbconv_edge_bwd_tc: null descriptor");
  GRL_REQUIRE(d->n_dst > 0 && d->n_src > 0 && d->n_edges >= 0, GRL_EINVAL, "grl_fbconv_edge_bwd_tc: bad sizes");
  GRL_REQUIRE(d->rowptr_src && d->x_src && d->wk && d->grad_x1 && d->grad_x_src && d->edge_grad_partials &&
                  (d->n_edges == 0 || (d->src_eid && d->edge_src && d->edge_dst && d->basis_bf16 && d->grad_basis_bf16)),
              GRL_EINVAL, "grl_fbconv_edge_bwd_tc: null pointer");
  GRL_REQUIRE(d->n_partials_edge > 0, GRL_EINVAL, "grl_fbconv_edge_bwd_tc: n_partials_edge must be > 0");
  GRL_REQUIRE(d->accumulate_grad_basis >= 0 && d->accumulate_grad_basis <= 2 &&
                  (d->accumulate_grad_basis != 2 || d->grad_basis_acc_mask), GRL_EINVAL,
              "grl_fbconv_edge_bwd_tc: accumulate_grad_basis=%d (2 needs grad_basis_acc_mask)", d->accumulate_grad_basis);
  const int smem2 = (int)sizeof(grl::EdgeBwd2Smem);
  if (grl::ensure_dynamic_smem((const void*)grl::fbconv_edge_bwd_tc2_kernel, smem2) != GRL_OK) return GRL_ECUDA;
  grl::fbconv_edge_bwd_tc2_kernel<<<d->n_partials_edge, grl::kThreads, smem2, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_fbconv_edge_bwd_tc");
}

int grl_edge_basis_bwd_tc(const GrlBasisDesc* d, grl_stream_t stream) {
  GRL_REQUIRE(d, GRL_EINVAL, "grl_edge_basis_bwd_tc: null descriptor");
  GRL_REQUIRE(d->n_edges > 0 && (d->dim == 2 || d->dim == 3), GRL_EINVAL, "grl_edge_basis_bwd_tc: n_edges=%d dim=%d",
              d->n_edges, d->dim);
  GRL_REQUIRE(d->edge_src && d->edge_dst && d->pos_src && d->pos_dst && d->ori && d->w1t && d->b1 && d->w2t && d->b2 &&
                  d->grad_basis_bf16 && d->grad_partials, GRL_EINVAL, "grl_edge_basis_bwd_tc: null pointer");
  GRL_REQUIRE(d->n_partials > 0, GRL_EINVAL, "grl_edge_basis_bwd_tc: n_partials must be > 0");
  const int smem = (int)sizeof(grl::BasisBwdTcSmem);
  if (grl::ensure_dynamic_smem((const void*)grl::edge_basis_bwd_tc_kernel, smem) != GRL_OK) return GRL_ECUDA;
  grl::edge_basis_bwd_tc_kernel<<<d->n_partials, grl::kThreads, smem, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_edge_basis_bwd_tc");
}

}  // extern "C"
