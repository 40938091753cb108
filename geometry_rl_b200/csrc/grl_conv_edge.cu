// Spatial half of the separable fibre-bundle convolution, forward and backward.
//   forward : x1[d] = sum_{e in in(d)} (basis[e] Wk^T) * x_src[src(e)]          (dst-sorted CSR order)
//   backward: g_x_src[s] = init[s] + sum_{e in out(s)} g_x1[dst(e)] * (basis[e] Wk^T)   (src-sorted order)
//             g_basis[e] = (g_x1[dst(e)] * x_src[src(e)]) Wk ;  gWk += (g_x1*x_src)^T basis[e]
// Reference: geometry_rl/modules/pyg_models/ponita/conv.py:71-87,116-149 (kernel Linear, message =
// kernel * x_j, torch_scatter.scatter sum over edge_index[1]); ponita/ponita.py:152-161 for EMPN.
// The segmented sums are sequential adds in CSR order inside one thread (no atomics): the result is
// bit-reproducible and follows the summation order of a sequential scatter over the coalesced COO.
#include "grl_common.cuh"

namespace grl {

struct EdgeFwdSmem {
  float BZ[kTileFloats];
  float XS[kTileFloats];
  float WkT[kWFloats];
  int src[kTE];
  int dst[kTE];
};

struct EdgeBwdSmem {
  float BZ[kTileFloats];
  float GX[kTileFloats];   // g_x1 rows gathered by dst
  float XS[kTileFloats];   // x_src rows gathered by src
  float GK[kTileFloats];   // g_kern = g_msg * x_src
  float WkT[kWFloats];
  float Wk[kWFloats];
  int eid[kTE];
  int src[kTE];
  int dst[kTE];
};

constexpr int kNodesPerBlock = 16;

__global__ void __launch_bounds__(kThreads) fbconv_edge_fwd_kernel(const GrlConvDesc d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EdgeFwdSmem& s = *reinterpret_cast<EdgeFwdSmem*>(smem_raw);
  const int tid = threadIdx.x, o = tid >> 4, cg = tid & 15;
  for (int i = tid; i < kWFloats; i += kThreads) s.WkT[i] = d.wk_t[i];
  const int n_blocks = (d.n_dst + kNodesPerBlock - 1) / kNodesPerBlock;
  for (int nb = blockIdx.x; nb < n_blocks; nb += gridDim.x) {
    const int n0 = nb * kNodesPerBlock;
    const int n1 = min(n0 + kNodesPerBlock, d.n_dst);
    const int p0 = d.rowptr_dst[n0], p1 = d.rowptr_dst[n1];
    int cur = n0;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = p0; base < p1; base += kTE) {
      const int cnt = min(kTE, p1 - base);
      __syncthreads();  // previous tile fully consumed (also orders the WkT fill on the first pass)
      if (tid < kTE) {
        s.src[tid] = (tid < cnt) ? d.edge_src[base + tid] : 0;
        s.dst[tid] = (tid < cnt) ? d.edge_dst[base + tid] : 0;
      }
      stage_rows_contig(s.BZ, d.basis + (size_t)base * kRow, cnt);
      __syncthreads();  // s.src visible for the gather
      stage_rows_gather(s.XS, d.x_src, s.src, cnt);
      cp_async_commit();
      cp_async_wait_all();
      __syncthreads();
      float acc[kTE][4];
      zero_acc(acc);
      gemm_tile<64>(s.BZ, kLDT, s.WkT, o, cg, acc);
#pragma unroll
      for (int j = 0; j < kTE; ++j) {
        if (j < cnt) {
          const int dn = s.dst[j];
          while (cur < dn) {  // flush finished node, zero-fill nodes without in-edges
            st4(d.x1 + (size_t)cur * kRow + o * kC + 4 * cg, sum);
            sum = make_float4(0.f, 0.f, 0.f, 0.f);
            ++cur;
          }
          const float4 xs = ld4(s.XS + (16 * j + o) * kLDT + 4 * cg);
          sum.x = fmaf(acc[j][0], xs.x, sum.x);
          sum.y = fmaf(acc[j][1], xs.y, sum.y);
          sum.z = fmaf(acc[j][2], xs.z, sum.z);
          sum.w = fmaf(acc[j][3], xs.w, sum.w);
        }
      }
    }
    while (cur < n1) {
      st4(d.x1 + (size_t)cur * kRow + o * kC + 4 * cg, sum);
      sum = make_float4(0.f, 0.f, 0.f, 0.f);
      ++cur;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) fbconv_edge_bwd_kernel(const GrlConvDesc d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EdgeBwdSmem& s = *reinterpret_cast<EdgeBwdSmem*>(smem_raw);
  const int tid = threadIdx.x, o = tid >> 4, cg = tid & 15;
  const int ni = tid >> 4, mi = tid & 15;
  for (int i = tid; i < kWFloats; i += kThreads) { s.WkT[i] = d.wk_t[i]; s.Wk[i] = d.wk[i]; }
  float gWk[4][4], unused[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) gWk[i][j] = 0.f;

  const int n_blocks = (d.n_src + kNodesPerBlock - 1) / kNodesPerBlock;
  for (int nb = blockIdx.x; nb < n_blocks; nb += gridDim.x) {
    const int n0 = nb * kNodesPerBlock;
    const int n1 = min(n0 + kNodesPerBlock, d.n_src);
    const int q0 = d.rowptr_src[n0], q1 = d.rowptr_src[n1];
    int cur = n0;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    auto flush = [&](int node) {
      const size_t off = (size_t)node * kRow + o * kC + 4 * cg;
      if (d.grad_x_src_init) {
        const float4 b = ldg4(d.grad_x_src_init + off);
        sum.x += b.x; sum.y += b.y; sum.z += b.z; sum.w += b.w;
      }
      st4(d.grad_x_src + off, sum);
      sum = make_float4(0.f, 0.f, 0.f, 0.f);
    };
    for (int base = q0; base < q1; base += kTE) {
      const int cnt = min(kTE, q1 - base);
      __syncthreads();
      if (tid < kTE) {
        const int e = (tid < cnt) ? d.src_eid[base + tid] : 0;
        s.eid[tid] = e;
        s.src[tid] = (tid < cnt) ? d.edge_src[e] : 0;
        s.dst[tid] = (tid < cnt) ? d.edge_dst[e] : 0;
      }
      __syncthreads();
      stage_rows_gather(s.BZ, d.basis, s.eid, cnt);
      stage_rows_gather(s.GX, d.grad_x1, s.dst, cnt);
      stage_rows_gather(s.XS, d.x_src, s.src, cnt);
      cp_async_commit();
      cp_async_wait_all();
      __syncthreads();
      float acc[kTE][4];
      zero_acc(acc);
      gemm_tile<64>(s.BZ, kLDT, s.WkT, o, cg, acc);  // kern
#pragma unroll
      for (int j = 0; j < kTE; ++j) {
        const int off = (16 * j + o) * kLDT + 4 * cg;
        float4 gk = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < cnt) {
          const int sn = s.src[j];
          while (cur < sn) { flush(cur); ++cur; }
          const float4 gm = ld4(s.GX + off);
          const float4 xs = ld4(s.XS + off);
          sum.x = fmaf(gm.x, acc[j][0], sum.x);
          sum.y = fmaf(gm.y, acc[j][1], sum.y);
          sum.z = fmaf(gm.z, acc[j][2], sum.z);
          sum.w = fmaf(gm.w, acc[j][3], sum.w);
          gk = make_float4(gm.x * xs.x, gm.y * xs.y, gm.z * xs.z, gm.w * xs.w);
        }
        st4(s.GK + off, gk);
      }
      __syncthreads();
      // g_basis[e] = GK . Wk   (B[k = c][j] = Wk[c][j])
      zero_acc(acc);
      gemm_tile<64>(s.GK, kLDT, s.Wk, o, cg, acc);
#pragma unroll
      for (int j = 0; j < kTE; ++j) {
        if (j < cnt) {
          float* dst = d.grad_basis + (size_t)s.eid[j] * kRow + o * kC + 4 * cg;
          float4 v = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
          if (d.accumulate_grad_basis) {
            const float4 old = ld4(dst);
            v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
          }
          st4(dst, v);
        }
      }
      // gWk[c][j] += sum_r GK[r][c] BZ[r][j]
      wgrad_tile<false>(s.GK, kLDT, s.BZ, kLDT, ni, mi, gWk, unused);
    }
    while (cur < n1) { flush(cur); ++cur; }
  }
  float* P = d.edge_grad_partials + (size_t)blockIdx.x * kWFloats;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    st4(P + (4 * ni + i) * kC + 4 * mi, make_float4(gWk[i][0], gWk[i][1], gWk[i][2], gWk[i][3]));
}

}  // namespace grl

extern "C" {

int grl_fbconv_edge_fwd(const GrlConvDesc* d, grl_stream_t stream) {
  GRL_REQUIRE(d, GRL_EINVAL, "grl_fbconv_edge_fwd: null descriptor");
  GRL_REQUIRE(d->n_dst > 0 && d->n_src > 0 && d->n_edges >= 0, GRL_EINVAL, "grl_fbconv_edge_fwd: n_src=%d n_dst=%d E=%d",
              d->n_src, d->n_dst, d->n_edges);
  GRL_REQUIRE(d->rowptr_dst && d->x_src && d->wk_t && d->x1 && (d->n_edges == 0 || (d->edge_src && d->edge_dst && d->basis)),
              GRL_EINVAL, "grl_fbconv_edge_fwd: null pointer");
  const int smem = (int)sizeof(grl::EdgeFwdSmem);
  if (grl::ensure_dynamic_smem((const void*)grl::fbconv_edge_fwd_kernel, smem) != GRL_OK) return GRL_ECUDA;
  const int n_blocks = (d->n_dst + grl::kNodesPerBlock - 1) / grl::kNodesPerBlock;
  int grid = 2 * grl::sm_count();
  if (grid > n_blocks) grid = n_blocks;
  grl::fbconv_edge_fwd_kernel<<<grid, grl::kThreads, smem, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_fbconv_edge_fwd");
}

int grl_fbconv_edge_bwd(const GrlConvDesc* d, grl_stream_t stream) {
  GRL_REQUIRE(d, GRL_EINVAL, "grl_fbconv_edge_bwd: null descriptor");
  GRL_REQUIRE(d->n_dst > 0 && d->n_src > 0 && d->n_edges >= 0, GRL_EINVAL, "grl_fbconv_edge_bwd: bad sizes");
  GRL_REQUIRE(d->rowptr_src && d->x_src && d->wk_t && d->wk && d->grad_x1 && d->grad_x_src && d->edge_grad_partials &&
                  (d->n_edges == 0 || (d->src_eid && d->edge_src && d->edge_dst && d->basis && d->grad_basis)),
              GRL_EINVAL, "grl_fbconv_edge_bwd: null pointer");
  GRL_REQUIRE(d->n_partials_edge > 0, GRL_EINVAL, "grl_fbconv_edge_bwd: n_partials_edge must be > 0");
  const int smem = (int)sizeof(grl::EdgeBwdSmem);
  if (grl::ensure_dynamic_smem((const void*)grl::fbconv_edge_bwd_kernel, smem) != GRL_OK) return GRL_ECUDA;
  grl::fbconv_edge_bwd_kernel<<<d->n_partials_edge, grl::kThreads, smem, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_fbconv_edge_bwd");
}

}  // extern "C"
