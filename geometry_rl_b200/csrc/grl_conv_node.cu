// Node half of the separable fibre-bundle convolution + ConvNeXt update, forward and backward.
//   x2[n][p][c] = 1/16 sum_o x1[n][o][c] fk[o][p][c] + bias[c]
//   y = LayerNorm_c(x2);  h = GELU(y W1^T + b1);  out = x_dst + h W2^T + b2
// Reference: geometry_rl/modules/pyg_models/ponita/conv.py:88-114 (fibre einsum "boc,opc->bpc" / 16,
// bias, node_mlp = LayerNorm -> Linear(64,256) -> GELU -> Linear(256,64), residual) and
// ponita/ponita.py:163-175,219-230 for the EMPN layer (same math, fk passed pre-transposed).
// The hidden dimension (256) is processed in four 64-wide chunks so that every contraction is a
// [128 x 64] x [64 x 64] tile GEMM; weight chunks are streamed with cp.async, double buffered.
#include "grl_common.cuh"

namespace grl {

struct NodeFwdSmem {
  float XH[kTileFloats];  // x1 tile, later the hidden chunk
  float Y[kTileFloats];
  float B0[kWFloats];
  float B1[kWFloats];
  float b1[kH];
  float b2[kC], bias[kC], lng[kC], lnb[kC];
};

struct NodeBwdSmem {
  float Y[kTileFloats];
  float XH[kTileFloats];  // x-hat
  float GZ[kTileFloats];  // grad_out tile
  float HS[kTileFloats];  // x1 tile -> hidden chunk -> column-sum scratch
  float GP[kTileFloats];  // g_pre chunk -> g_x2
  float B0[kWFloats];
  float B1[kWFloats];
  float b1[kH];
  float lng[kC], lnb[kC], bias[kC];
  float rstd[kTM];
  float acc_gb1[kH];
  float acc_gb2[kC], acc_glng[kC], acc_glnb[kC], acc_gbias[kC];
};

// offsets inside one node partial slot (floats)
constexpr int kOffGW1 = 0;
constexpr int kOffGB1 = kOffGW1 + kH * kC;
constexpr int kOffGW2 = kOffGB1 + kH;
constexpr int kOffGB2 = kOffGW2 + kC * kH;
constexpr int kOffGLNG = kOffGB2 + kC;
constexpr int kOffGLNB = kOffGLNG + kC;
constexpr int kOffGBIAS = kOffGLNB + kC;
constexpr int kOffGFK = kOffGBIAS + kC;
static_assert(kOffGFK + kO * kO * kC == GRL_NODE_GRAD_FLOATS, "partial layout");

// Fibre convolution + bias + LayerNorm for the 8 rows (16 j + p) this thread owns.
// X: x1 tile in smem; fk[o] = fiber_kernel[o][p][4cg..]. Outputs y (affine) and xhat, rstd per row.
__device__ __forceinline__ void fiber_ln(const float* __restrict__ X, const float4 (&fk)[kO], int cg, float4 bias,
                                         float4 g, float4 b, float4 (&y)[kTE], float4 (&xh)[kTE], float (&rs)[kTE]) {
#pragma unroll
  for (int j = 0; j < kTE; ++j) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int o = 0; o < kO; ++o) {
      const float4 x = ld4(X + (16 * j + o) * kLDT + 4 * cg);
      a.x = fmaf(x.x, fk[o].x, a.x);
      a.y = fmaf(x.y, fk[o].y, a.y);
      a.z = fmaf(x.z, fk[o].z, a.z);
      a.w = fmaf(x.w, fk[o].w, a.w);
    }
    a.x = a.x * 0.0625f + bias.x;
    a.y = a.y * 0.0625f + bias.y;
    a.z = a.z * 0.0625f + bias.z;
    a.w = a.w * 0.0625f + bias.w;
    const float mean = row_sum16((a.x + a.y) + (a.z + a.w)) * (1.0f / 64.0f);
    const float dx = a.x - mean, dy = a.y - mean, dz = a.z - mean, dw = a.w - mean;
    const float var = row_sum16((dx * dx + dy * dy) + (dz * dz + dw * dw)) * (1.0f / 64.0f);
    const float r = rsqrtf(var + 1e-5f);
    rs[j] = r;
    xh[j] = make_float4(dx * r, dy * r, dz * r, dw * r);
    y[j] = make_float4(xh[j].x * g.x + b.x, xh[j].y * g.y + b.y, xh[j].z * g.z + b.z, xh[j].w * g.w + b.w);
  }
}

__global__ void __launch_bounds__(kThreads, 1) fbconv_node_fwd_kernel(const GrlConvDesc d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NodeFwdSmem& s = *reinterpret_cast<NodeFwdSmem*>(smem_raw);
  const int tid = threadIdx.x, o = tid >> 4, cg = tid & 15;  // o doubles as the output orientation p
  for (int i = tid; i < kH; i += kThreads) s.b1[i] = d.b1[i];
  if (tid < kC) { s.b2[tid] = d.b2[tid]; s.bias[tid] = d.bias[tid]; s.lng[tid] = d.ln_g[tid]; s.lnb[tid] = d.ln_b[tid]; }
  float4 fk[kO];
#pragma unroll
  for (int oo = 0; oo < kO; ++oo) fk[oo] = ldg4(d.fiber_kernel + ((size_t)(oo * kO + o)) * kC + 4 * cg);

  const int n_tiles = (d.n_dst + kTE - 1) / kTE;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int n0 = tile * kTE, cnt = min(kTE, d.n_dst - n0);
    __syncthreads();
    stage_rows_contig(s.XH, d.x1 + (size_t)n0 * kRow, cnt);
    stage_w64(s.B0, d.w1_t);
    cp_async_commit();
    stage_w64(s.B1, d.w2_t);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    {
      float4 y[kTE], xh[kTE];
      float rs[kTE];
      fiber_ln(s.XH, fk, cg, ld4(s.bias + 4 * cg), ld4(s.lng + 4 * cg), ld4(s.lnb + 4 * cg), y, xh, rs);
#pragma unroll
      for (int j = 0; j < kTE; ++j) st4(s.Y + (16 * j + o) * kLDT + 4 * cg, y[j]);
    }
    __syncthreads();  // Y visible, x1 tile no longer needed
    float z[kTE][4];
    zero_acc(z);
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      if (q > 0) {
        cp_async_wait<1>();
        __syncthreads();
      }
      float acc[kTE][4];
      const float4 bq = ld4(s.b1 + q * kC + 4 * cg);
#pragma unroll
      for (int j = 0; j < kTE; ++j) { acc[j][0] = bq.x; acc[j][1] = bq.y; acc[j][2] = bq.z; acc[j][3] = bq.w; }
      gemm_tile<64>(s.Y, kLDT, s.B0, o, cg, acc);
#pragma unroll
      for (int j = 0; j < kTE; ++j)
        st4(s.XH + (16 * j + o) * kLDT + 4 * cg,
            make_float4(gelu_f(acc[j][0]), gelu_f(acc[j][1]), gelu_f(acc[j][2]), gelu_f(acc[j][3])));
      cp_async_wait_all();
      __syncthreads();  // W2T_q landed, hidden chunk visible, everyone done with B0
      if (q < 3) {
        stage_w64(s.B0, d.w1_t + (size_t)(q + 1) * kWFloats);
        cp_async_commit();
      }
      gemm_tile<64>(s.XH, kLDT, s.B1, o, cg, z);
      __syncthreads();  // everyone done with B1 and the hidden chunk
      if (q < 3) {
        stage_w64(s.B1, d.w2_t + (size_t)(q + 1) * kWFloats);
        cp_async_commit();
      }
    }
    const float4 b2v = ld4(s.b2 + 4 * cg);
#pragma unroll
    for (int j = 0; j < kTE; ++j) {
      if (j < cnt) {
        const size_t off = (size_t)(n0 + j) * kRow + o * kC + 4 * cg;
        const float4 xd = ldg4(d.x_dst + off);
        float4 v = make_float4(xd.x + (z[j][0] + b2v.x), xd.y + (z[j][1] + b2v.y), xd.z + (z[j][2] + b2v.z),
                               xd.w + (z[j][3] + b2v.w));
        if (d.accumulate_out) {
          const float4 old = ld4(d.out + off);
          v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
        }
        st4(d.out + off, v);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) fbconv_node_bwd_kernel(const GrlConvDesc d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NodeBwdSmem& s = *reinterpret_cast<NodeBwdSmem*>(smem_raw);
  const int tid = threadIdx.x, o = tid >> 4, cg = tid & 15;
  const int ni = tid >> 4, mi = tid & 15;
  float* P = d.node_grad_partials + (size_t)blockIdx.x * GRL_NODE_GRAD_FLOATS;
  for (int i = tid; i < GRL_NODE_GRAD_FLOATS; i += kThreads) P[i] = 0.f;
  for (int i = tid; i < kH; i += kThreads) { s.b1[i] = d.b1[i]; s.acc_gb1[i] = 0.f; }
  if (tid < kC) {
    s.bias[tid] = d.bias[tid]; s.lng[tid] = d.ln_g[tid]; s.lnb[tid] = d.ln_b[tid];
    s.acc_gb2[tid] = 0.f; s.acc_glng[tid] = 0.f; s.acc_glnb[tid] = 0.f; s.acc_gbias[tid] = 0.f;
  }

  const int n_tiles = (d.n_dst + kTE - 1) / kTE;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int n0 = tile * kTE, cnt = min(kTE, d.n_dst - n0);
    __syncthreads();
    stage_rows_contig(s.HS, d.x1 + (size_t)n0 * kRow, cnt);
    stage_rows_contig(s.GZ, d.grad_out + (size_t)n0 * kRow, cnt);
    stage_w64(s.B0, d.w1_t);
    cp_async_commit();
    stage_w64(s.B1, d.w2_c);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    {  // recompute x2 -> LayerNorm
      float4 fk[kO];
#pragma unroll
      for (int oo = 0; oo < kO; ++oo) fk[oo] = ldg4(d.fiber_kernel + ((size_t)(oo * kO + o)) * kC + 4 * cg);
      float4 y[kTE], xh[kTE];
      float rs[kTE];
      fiber_ln(s.HS, fk, cg, ld4(s.bias + 4 * cg), ld4(s.lng + 4 * cg), ld4(s.lnb + 4 * cg), y, xh, rs);
#pragma unroll
      for (int j = 0; j < kTE; ++j) {
        st4(s.Y + (16 * j + o) * kLDT + 4 * cg, y[j]);
        st4(s.XH + (16 * j + o) * kLDT + 4 * cg, xh[j]);
        if (cg == 0) s.rstd[16 * j + o] = rs[j];
      }
    }
    __syncthreads();

    float gy[kTE][4];
    zero_acc(gy);
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      if (q > 0) {
        cp_async_wait_all();
        __syncthreads();  // B0 = W1T_q
      }
      float dG[kTE][4];
      {
        float acc[kTE][4];
        const float4 bq = ld4(s.b1 + q * kC + 4 * cg);
#pragma unroll
        for (int j = 0; j < kTE; ++j) { acc[j][0] = bq.x; acc[j][1] = bq.y; acc[j][2] = bq.z; acc[j][3] = bq.w; }
        gemm_tile<64>(s.Y, kLDT, s.B0, o, cg, acc);
#pragma unroll
        for (int j = 0; j < kTE; ++j) {
#pragma unroll
          for (int c = 0; c < 4; ++c) dG[j][c] = gelu_grad_f(acc[j][c]);
          st4(s.HS + (16 * j + o) * kLDT + 4 * cg,
              make_float4(gelu_f(acc[j][0]), gelu_f(acc[j][1]), gelu_f(acc[j][2]), gelu_f(acc[j][3])));
        }
      }
      cp_async_wait_all();
      __syncthreads();  // B1 = W2C_q landed; hidden chunk visible; everyone done with B0
      stage_w64(s.B0, d.w1 + (size_t)q * kWFloats);  // rows q*64.. of W1: B[k = n'][m] = W1[q*64+n'][m]
      cp_async_commit();
      {
        float acc[kTE][4];
        zero_acc(acc);
        gemm_tile<64>(s.GZ, kLDT, s.B1, o, cg, acc);  // g_h chunk
#pragma unroll
        for (int j = 0; j < kTE; ++j)
          st4(s.GP + (16 * j + o) * kLDT + 4 * cg,
              make_float4(acc[j][0] * dG[j][0], acc[j][1] * dG[j][1], acc[j][2] * dG[j][2], acc[j][3] * dG[j][3]));
      }
      __syncthreads();  // g_pre chunk visible; everyone done with B1
      if (q < 3) {
        stage_w64(s.B1, d.w2_c + (size_t)(q + 1) * kWFloats);
        cp_async_commit();
      }
      {  // gW2[n][q*64+m] += sum_r GZ[r][n] HS[r][m]  (+ gb2 once)
        float g[4][4], cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) g[i][j] = 0.f;
        if (q == 0) wgrad_tile<true>(s.GZ, kLDT, s.HS, kLDT, ni, mi, g, cs);
        else wgrad_tile<false>(s.GZ, kLDT, s.HS, kLDT, ni, mi, g, cs);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float* p = P + kOffGW2 + (size_t)(4 * ni + i) * kH + q * kC + 4 * mi;
          const float4 old = ld4(p);
          st4(p, make_float4(old.x + g[i][0], old.y + g[i][1], old.z + g[i][2], old.w + g[i][3]));
        }
        if (q == 0 && mi == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) s.acc_gb2[4 * ni + i] += cs[i];
        }
      }
      {  // gW1[q*64+n'][k] += sum_r GP[r][n'] Y[r][k];  gb1[q*64+n'] += colsum
        float g[4][4], cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) g[i][j] = 0.f;
        wgrad_tile<true>(s.GP, kLDT, s.Y, kLDT, ni, mi, g, cs);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float* p = P + kOffGW1 + (size_t)(q * kC + 4 * ni + i) * kC + 4 * mi;
          const float4 old = ld4(p);
          st4(p, make_float4(old.x + g[i][0], old.y + g[i][1], old.z + g[i][2], old.w + g[i][3]));
        }
        if (mi == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) s.acc_gb1[q * kC + 4 * ni + i] += cs[i];
        }
      }
      if (q < 3) cp_async_wait<1>(); else cp_async_wait_all();
      __syncthreads();  // B0 = W1 rows of chunk q
      gemm_tile<64>(s.GP, kLDT, s.B0, o, cg, gy);  // g_y += g_pre . W1_q
      __syncthreads();  // everyone done with B0, GP, HS
      if (q < 3) {
        stage_w64(s.B0, d.w1_t + (size_t)(q + 1) * kWFloats);
        cp_async_commit();
      }
    }

    // ---- LayerNorm backward, bias gradient ----------------------------------------------------
    {
      const float4 g4 = ld4(s.lng + 4 * cg);
      float4 pl_g = make_float4(0.f, 0.f, 0.f, 0.f), pl_b = pl_g, pl_bias = pl_g;
#pragma unroll
      for (int j = 0; j < kTE; ++j) {
        const int r = 16 * j + o;
        const float4 xh = ld4(s.XH + r * kLDT + 4 * cg);
        pl_g.x = fmaf(gy[j][0], xh.x, pl_g.x); pl_g.y = fmaf(gy[j][1], xh.y, pl_g.y);
        pl_g.z = fmaf(gy[j][2], xh.z, pl_g.z); pl_g.w = fmaf(gy[j][3], xh.w, pl_g.w);
        pl_b.x += gy[j][0]; pl_b.y += gy[j][1]; pl_b.z += gy[j][2]; pl_b.w += gy[j][3];
        const float hx = gy[j][0] * g4.x, hy = gy[j][1] * g4.y, hz = gy[j][2] * g4.z, hw = gy[j][3] * g4.w;
        const float m1 = row_sum16((hx + hy) + (hz + hw)) * (1.0f / 64.0f);
        const float m2 = row_sum16((hx * xh.x + hy * xh.y) + (hz * xh.z + hw * xh.w)) * (1.0f / 64.0f);
        const float rs = s.rstd[r];
        const float4 gx = make_float4(rs * (hx - m1 - xh.x * m2), rs * (hy - m1 - xh.y * m2),
                                      rs * (hz - m1 - xh.z * m2), rs * (hw - m1 - xh.w * m2));
        pl_bias.x += gx.x; pl_bias.y += gx.y; pl_bias.z += gx.z; pl_bias.w += gx.w;
        st4(s.GP + r * kLDT + 4 * cg, gx);
      }
      st4(s.HS + (0 * kO + o) * kC + 4 * cg, pl_g);
      st4(s.HS + (1 * kO + o) * kC + 4 * cg, pl_b);
      st4(s.HS + (2 * kO + o) * kC + 4 * cg, pl_bias);
    }
    __syncthreads();
    if (tid < 3 * kC) {
      const int kind = tid >> 6, c = tid & 63;
      float t = 0.f;
#pragma unroll
      for (int oo = 0; oo < kO; ++oo) t += s.HS[(kind * kO + oo) * kC + c];
      float* dst = kind == 0 ? s.acc_glng : (kind == 1 ? s.acc_glnb : s.acc_gbias);
      dst[c] += t;
    }
    // ---- fibre convolution backward -----------------------------------------------------------
    {
      float4 fko[kO];  // fiber_kernel[o][p][4cg..] for all p
#pragma unroll
      for (int p = 0; p < kO; ++p) fko[p] = ldg4(d.fiber_kernel + ((size_t)(o * kO + p)) * kC + 4 * cg);
      float4 x1r[kTE];
#pragma unroll
      for (int j = 0; j < kTE; ++j) {
        x1r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < cnt) {
          x1r[j] = ldg4(d.x1 + (size_t)(n0 + j) * kRow + o * kC + 4 * cg);
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int p = 0; p < kO; ++p) {
            const float4 g = ld4(s.GP + (16 * j + p) * kLDT + 4 * cg);
            a.x = fmaf(g.x, fko[p].x, a.x);
            a.y = fmaf(g.y, fko[p].y, a.y);
            a.z = fmaf(g.z, fko[p].z, a.z);
            a.w = fmaf(g.w, fko[p].w, a.w);
          }
          st4(d.grad_x1 + (size_t)(n0 + j) * kRow + o * kC + 4 * cg,
              make_float4(a.x * 0.0625f, a.y * 0.0625f, a.z * 0.0625f, a.w * 0.0625f));
        }
      }
      // g_fk[o][p][c] += 1/16 sum_j x1[j][o][c] g_x2[j][p][c]
#pragma unroll
      for (int p = 0; p < kO; ++p) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < kTE; ++j) {
          const float4 g = ld4(s.GP + (16 * j + p) * kLDT + 4 * cg);
          t.x = fmaf(x1r[j].x, g.x, t.x);
          t.y = fmaf(x1r[j].y, g.y, t.y);
          t.z = fmaf(x1r[j].z, g.z, t.z);
          t.w = fmaf(x1r[j].w, g.w, t.w);
        }
        float* pf = P + kOffGFK + ((size_t)(o * kO + p)) * kC + 4 * cg;
        const float4 old = ld4(pf);
        st4(pf, make_float4(old.x + t.x * 0.0625f, old.y + t.y * 0.0625f, old.z + t.z * 0.0625f,
                            old.w + t.w * 0.0625f));
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < kH; i += kThreads) P[kOffGB1 + i] = s.acc_gb1[i];
  if (tid < kC) {
    P[kOffGB2 + tid] = s.acc_gb2[tid];
    P[kOffGLNG + tid] = s.acc_glng[tid];
    P[kOffGLNB + tid] = s.acc_glnb[tid];
    P[kOffGBIAS + tid] = s.acc_gbias[tid];
  }
}

}  // namespace grl

extern "C" {

static int check_node_desc(const GrlConvDesc* d, const char* who, bool bwd) {
  GRL_REQUIRE(d, GRL_EINVAL, "%s: null descriptor", who);
  GRL_REQUIRE(d->n_dst > 0, GRL_EINVAL, "%s: n_dst=%d", who, d->n_dst);
  GRL_REQUIRE(d->x1 && d->fiber_kernel && d->bias && d->ln_g && d->ln_b && d->w1_t && d->b1, GRL_EINVAL,
              "%s: null pointer", who);
  if (bwd) {
    GRL_REQUIRE(d->w1 && d->w2_c && d->grad_out && d->grad_x1 && d->node_grad_partials && d->n_partials_node > 0,
                GRL_EINVAL, "%s: null backward pointer", who);
  } else {
    GRL_REQUIRE(d->w2_t && d->b2 && d->x_dst && d->out, GRL_EINVAL, "%s: null forward pointer", who);
  }
  return GRL_OK;
}

int grl_fbconv_node_fwd(const GrlConvDesc* d, grl_stream_t stream) {
  const int rc = check_node_desc(d, "grl_fbconv_node_fwd", false);
  if (rc != GRL_OK) return rc;
  const int smem = (int)sizeof(grl::NodeFwdSmem);
  if (grl::ensure_dynamic_smem((const void*)grl::fbconv_node_fwd_kernel, smem) != GRL_OK) return GRL_ECUDA;
  const int n_tiles = (d->n_dst + grl::kTE - 1) / grl::kTE;
  int grid = 2 * grl::sm_count();
  if (grid > n_tiles) grid = n_tiles;
  grl::fbconv_node_fwd_kernel<<<grid, grl::kThreads, smem, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_fbconv_node_fwd");
}

int grl_fbconv_node_bwd(const GrlConvDesc* d, grl_stream_t stream) {
  const int rc = check_node_desc(d, "grl_fbconv_node_bwd", true);
  if (rc != GRL_OK) return rc;
  const int smem = (int)sizeof(grl::NodeBwdSmem);
  if (grl::ensure_dynamic_smem((const void*)grl::fbconv_node_bwd_kernel, smem) != GRL_OK) return GRL_ECUDA;
  grl::fbconv_node_bwd_kernel<<<d->n_partials_node, grl::kThreads, smem, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_fbconv_node_bwd");
}

}  // extern "C"
