// Edge basis: invariants -> polynomial features -> Linear(14,64) GELU Linear(64,64) GELU, forward and
// backward (weight gradients only; positions carry no gradient in the reference either).
// Reference: geometry_rl/modules/pyg_models/hepi.py:109-123 (compute_invariants), :76-82 (basis_fn),
// ponita/ponita.py:233-244 (PolynomialFeatures), ponita/ponita.py:327-347 (same for EMPN).
#include "grl_common.cuh"

namespace grl {

constexpr int kLDF = 20;  // padded stride of the [128][16] feature tile

struct BasisSmemFwd {
  float F[kTM * kLDF];
  float H1[kTileFloats];
  float W1t[16 * kC];
  float W2t[kWFloats];
  float b1[kC];
  float b2[kC];
};

struct BasisSmemBwd {
  float F[kTM * kLDF];
  float P1[kTileFloats];   // pre-activation 1, later g_pre1
  float H1[kTileFloats];
  float GP2[kTileFloats];  // g_pre2
  float W1t[16 * kC];
  float W2t[kWFloats];
  float W2[kWFloats];
  float b1[kC];
  float b2[kC];
  float red[kO * kC];      // cross-warp column sums
};

// One thread per tile row computes the 14 invariant features of (edge, orientation).
__device__ __forceinline__ void basis_features(const GrlBasisDesc& d, int tile, float* __restrict__ F) {
  const int r = threadIdx.x;
  if (r < kTM) {
    const int e = tile * kTE + (r >> 4), o = r & 15;
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = 0.f;
    if (e < d.n_edges) {
      const float* ps = d.pos_src + 3 * (size_t)d.edge_src[e];
      const float* pd = d.pos_dst + 3 * (size_t)d.edge_dst[e];
      const float rx = ps[0] - pd[0], ry = ps[1] - pd[1], rz = (d.dim == 3) ? ps[2] - pd[2] : 0.f;
      const float ox = d.ori[3 * o], oy = d.ori[3 * o + 1], oz = (d.dim == 3) ? d.ori[3 * o + 2] : 0.f;
      const float i1 = (rx * ox + ry * oy) + rz * oz;
      const float tx = rx - i1 * ox, ty = ry - i1 * oy, tz = rz - i1 * oz;
      const float i2 = sqrtf((tx * tx + ty * ty) + tz * tz);
      f[0] = i1; f[1] = i2;
      f[2] = i1 * i1; f[3] = i1 * i2; f[4] = i2 * i1; f[5] = i2 * i2;
      f[6] = f[2] * i1; f[7] = f[2] * i2; f[8] = f[3] * i1; f[9] = f[3] * i2;
      f[10] = f[4] * i1; f[11] = f[4] * i2; f[12] = f[5] * i1; f[13] = f[5] * i2;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) st4(F + r * kLDF + 4 * i, make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]));
  }
}

__global__ void __launch_bounds__(kThreads) edge_basis_fwd_kernel(const GrlBasisDesc d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BasisSmemFwd& s = *reinterpret_cast<BasisSmemFwd*>(smem_raw);
  const int tid = threadIdx.x, o = tid >> 4, cg = tid & 15;
  for (int i = tid; i < 16 * kC; i += kThreads) s.W1t[i] = d.w1t[i];
  for (int i = tid; i < kWFloats; i += kThreads) s.W2t[i] = d.w2t[i];
  if (tid < kC) { s.b1[tid] = d.b1[tid]; s.b2[tid] = d.b2[tid]; }
  const int n_tiles = (d.n_edges + kTE - 1) / kTE;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();  // previous tile done with F/H1; weights visible on first pass
    basis_features(d, tile, s.F);
    __syncthreads();
    float acc[kTE][4];
    const float4 b1v = ld4(s.b1 + 4 * cg);
#pragma unroll
    for (int j = 0; j < kTE; ++j) { acc[j][0] = b1v.x; acc[j][1] = b1v.y; acc[j][2] = b1v.z; acc[j][3] = b1v.w; }
    gemm_tile<16>(s.F, kLDF, s.W1t, o, cg, acc);
#pragma unroll
    for (int j = 0; j < kTE; ++j)
      st4(s.H1 + (16 * j + o) * kLDT + 4 * cg,
          make_float4(gelu_f(acc[j][0]), gelu_f(acc[j][1]), gelu_f(acc[j][2]), gelu_f(acc[j][3])));
    __syncthreads();
    const float4 b2v = ld4(s.b2 + 4 * cg);
#pragma unroll
    for (int j = 0; j < kTE; ++j) { acc[j][0] = b2v.x; acc[j][1] = b2v.y; acc[j][2] = b2v.z; acc[j][3] = b2v.w; }
    gemm_tile<64>(s.H1, kLDT, s.W2t, o, cg, acc);
#pragma unroll
    for (int j = 0; j < kTE; ++j) {
      const int e = tile * kTE + j;
      if (e < d.n_edges)
        st4(d.basis + (size_t)e * kRow + o * kC + 4 * cg,
            make_float4(gelu_f(acc[j][0]), gelu_f(acc[j][1]), gelu_f(acc[j][2]), gelu_f(acc[j][3])));
    }
  }
}

// Backward: recompute the two hidden layers, then
//   g_pre2 = g_basis * gelu'(pre2);  gW2 += g_pre2^T H1;  gb2 += colsum(g_pre2)
//   g_H1 = g_pre2 W2;  g_pre1 = g_H1 * gelu'(pre1);  gW1 += g_pre1^T F;  gb1 += colsum(g_pre1)
// Weight-gradient accumulators live in registers for the whole (persistent) CTA and are written
// once to this CTA's partial slot; grl_reduce_partials sums the slots in fixed order.
__global__ void __launch_bounds__(kThreads, 1) edge_basis_bwd_kernel(const GrlBasisDesc d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BasisSmemBwd& s = *reinterpret_cast<BasisSmemBwd*>(smem_raw);
  const int tid = threadIdx.x, o = tid >> 4, cg = tid & 15;
  for (int i = tid; i < 16 * kC; i += kThreads) s.W1t[i] = d.w1t[i];
  for (int i = tid; i < kWFloats; i += kThreads) { s.W2t[i] = d.w2t[i]; s.W2[i] = d.w2[i]; }
  if (tid < kC) { s.b1[tid] = d.b1[tid]; s.b2[tid] = d.b2[tid]; }

  float gW2[4][4], gb2[4], gW1[4], gb1 = 0.f;  // gW2 block (4ni..,4mi..); gW1[n1][4 f1..] with n1 = tid>>2, f1 = tid&3
#pragma unroll
  for (int i = 0; i < 4; ++i) { gb2[i] = 0.f; gW1[i] = 0.f; for (int j = 0; j < 4; ++j) gW2[i][j] = 0.f; }
  const int ni = tid >> 4, mi = tid & 15;
  const int n1 = tid >> 2, f1 = tid & 3;

  const int n_tiles = (d.n_edges + kTE - 1) / kTE;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    basis_features(d, tile, s.F);
    __syncthreads();
    float acc[kTE][4];
    const float4 b1v = ld4(s.b1 + 4 * cg);
#pragma unroll
    for (int j = 0; j < kTE; ++j) { acc[j][0] = b1v.x; acc[j][1] = b1v.y; acc[j][2] = b1v.z; acc[j][3] = b1v.w; }
    gemm_tile<16>(s.F, kLDF, s.W1t, o, cg, acc);
#pragma unroll
    for (int j = 0; j < kTE; ++j) {
      const int off = (16 * j + o) * kLDT + 4 * cg;
      st4(s.P1 + off, make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]));
      st4(s.H1 + off, make_float4(gelu_f(acc[j][0]), gelu_f(acc[j][1]), gelu_f(acc[j][2]), gelu_f(acc[j][3])));
    }
    __syncthreads();
    const float4 b2v = ld4(s.b2 + 4 * cg);
#pragma unroll
    for (int j = 0; j < kTE; ++j) { acc[j][0] = b2v.x; acc[j][1] = b2v.y; acc[j][2] = b2v.z; acc[j][3] = b2v.w; }
    gemm_tile<64>(s.H1, kLDT, s.W2t, o, cg, acc);
#pragma unroll
    for (int j = 0; j < kTE; ++j) {
      const int e = tile * kTE + j;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < d.n_edges) g = ldg4(d.grad_basis + (size_t)e * kRow + o * kC + 4 * cg);
      st4(s.GP2 + (16 * j + o) * kLDT + 4 * cg,
          make_float4(g.x * gelu_grad_f(acc[j][0]), g.y * gelu_grad_f(acc[j][1]), g.z * gelu_grad_f(acc[j][2]),
                      g.w * gelu_grad_f(acc[j][3])));
    }
    __syncthreads();
    // gW2[n][k] += sum_r GP2[r][n] H1[r][k], gb2[n] += sum_r GP2[r][n]
    wgrad_tile<true>(s.GP2, kLDT, s.H1, kLDT, ni, mi, gW2, gb2);
    // g_H1 = GP2 . W2  (B[k = n][m] = W2[n][m])
    zero_acc(acc);
    gemm_tile<64>(s.GP2, kLDT, s.W2, o, cg, acc);
#pragma unroll
    for (int j = 0; j < kTE; ++j) {
      const int off = (16 * j + o) * kLDT + 4 * cg;
      const float4 p = ld4(s.P1 + off);
      st4(s.P1 + off, make_float4(acc[j][0] * gelu_grad_f(p.x), acc[j][1] * gelu_grad_f(p.y),
                                  acc[j][2] * gelu_grad_f(p.z), acc[j][3] * gelu_grad_f(p.w)));
    }
    __syncthreads();
    // gW1[n][f] += sum_r GP1[r][n] F[r][f];  gb1[n] += sum_r GP1[r][n]
#pragma unroll 4
    for (int r = 0; r < kTM; ++r) {
      const float a = s.P1[r * kLDT + n1];
      const float4 b = ld4(s.F + r * kLDF + 4 * f1);
      gW1[0] = fmaf(a, b.x, gW1[0]); gW1[1] = fmaf(a, b.y, gW1[1]);
      gW1[2] = fmaf(a, b.z, gW1[2]); gW1[3] = fmaf(a, b.w, gW1[3]);
      if (f1 == 0) gb1 += a;
    }
  }
  // write this CTA's partial slot: gW1[64][16] | gb1[64] | gW2[64][64] | gb2[64]
  float* P = d.grad_partials + (size_t)blockIdx.x * GRL_BASIS_GRAD_FLOATS;
  st4(P + n1 * 16 + 4 * f1, make_float4(gW1[0], gW1[1], gW1[2], gW1[3]));
  if (f1 == 0) P[64 * 16 + n1] = gb1;
  float* PW2 = P + 64 * 16 + 64;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    st4(PW2 + (4 * ni + i) * kC + 4 * mi, make_float4(gW2[i][0], gW2[i][1], gW2[i][2], gW2[i][3]));
  if (mi == 0) st4(PW2 + kWFloats + 4 * ni, make_float4(gb2[0], gb2[1], gb2[2], gb2[3]));
}

static int basis_grid(int n_edges, int max_ctas) {
  const int n_tiles = (n_edges + kTE - 1) / kTE;
  int g = n_tiles < max_ctas ? n_tiles : max_ctas;
  return g < 1 ? 1 : g;
}

}  // namespace grl

extern "C" {

int grl_edge_basis_fwd(const GrlBasisDesc* d, grl_stream_t stream) {
  GRL_REQUIRE(d, GRL_EINVAL, "grl_edge_basis_fwd: null descriptor");
  if (d->n_edges == 0) return GRL_OK;
  GRL_REQUIRE(d->n_edges > 0 && (d->dim == 2 || d->dim == 3), GRL_EINVAL, "grl_edge_basis_fwd: n_edges=%d dim=%d",
              d->n_edges, d->dim);
  GRL_REQUIRE(d->edge_src && d->edge_dst && d->pos_src && d->pos_dst && d->ori && d->w1t && d->b1 && d->w2t && d->b2 &&
                  d->basis, GRL_EINVAL, "grl_edge_basis_fwd: null pointer");
  const int smem = (int)sizeof(grl::BasisSmemFwd);
  if (grl::ensure_dynamic_smem((const void*)grl::edge_basis_fwd_kernel, smem) != GRL_OK) return GRL_ECUDA;
  const int grid = grl::basis_grid(d->n_edges, 3 * grl::sm_count());
  grl::edge_basis_fwd_kernel<<<grid, grl::kThreads, smem, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_edge_basis_fwd");
}

int grl_edge_basis_bwd(const GrlBasisDesc* d, grl_stream_t stream) {
  GRL_REQUIRE(d, GRL_EINVAL, "grl_edge_basis_bwd: null descriptor");
  GRL_REQUIRE(d->n_edges > 0 && (d->dim == 2 || d->dim == 3), GRL_EINVAL, "grl_edge_basis_bwd: n_edges=%d dim=%d",
              d->n_edges, d->dim);
  GRL_REQUIRE(d->edge_src && d->edge_dst && d->pos_src && d->pos_dst && d->ori && d->w1t && d->b1 && d->w2t && d->b2 &&
                  d->w2 && d->grad_basis && d->grad_partials, GRL_EINVAL, "grl_edge_basis_bwd: null pointer");
  GRL_REQUIRE(d->n_partials > 0, GRL_EINVAL, "grl_edge_basis_bwd: n_partials must be > 0");
  const int smem = (int)sizeof(grl::BasisSmemBwd);
  if (grl::ensure_dynamic_smem((const void*)grl::edge_basis_bwd_kernel, smem) != GRL_OK) return GRL_ECUDA;
  // every partial slot is written exactly once -> the grid IS the number of partial slots
  grl::edge_basis_bwd_kernel<<<d->n_partials, grl::kThreads, smem, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_edge_basis_bwd");
}

}  // extern "C"
