// K4: batched trust-region projection of a diagonal Gaussian policy, forward and backward.
// Replaces (all under geometry_rl/algorithms/trust_region_projections/):
//   projections/base_projection_layer.py:71-100   mean_projection (closed form, autograd through maha)
//   projections/kl_projection_layer.py:15-111     KL projection, diag branch; the ITPAL op
//       cpp_projection.BatchedDiagCovOnlyProjection (:162-204: numpy round trip, OpenMP + NLopt on the
//       CPU, float64) becomes an in-register safeguarded Newton solve of the 1-D dual, also float64
//   projections/w2_projection_layer.py:15-68      commuting W2 projection (scale_prec = True)
// "v" is the diagonal of the matrix the reference calls std (numerically the variance).
// One thread per sample, k <= GRL_MAX_ACTION_DIM values in registers.
#include "grl_common.cuh"

namespace grl {

constexpr int kMaxK = GRL_MAX_ACTION_DIM;

struct KlEval {
  double kl, dkl;
};

// KL(eta) of c~(eta) = (eta + 1) / (eta / o + 1 / c) against o, and d KL / d eta.
__device__ __forceinline__ KlEval kl_eval(const double (&c)[kMaxK], const double (&o)[kMaxK], int k, double eta) {
  KlEval r{0.0, 0.0};
#pragma unroll
  for (int i = 0; i < kMaxK; ++i) {
    if (i < k) {
      const double D = eta / o[i] + 1.0 / c[i];
      const double ct = (eta + 1.0) / D;
      r.kl += ct / o[i] - 1.0 + log(o[i]) - log(ct);
      const double dct = (1.0 / c[i] - 1.0 / o[i]) / (D * D);
      r.dkl += (1.0 / o[i] - 1.0 / ct) * dct;
    }
  }
  r.kl *= 0.5;
  r.dkl *= 0.5;
  return r;
}

__global__ void __launch_bounds__(128) trpl_fwd_kernel(const GrlProjDesc d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.batch) return;
  const int k = d.k;
  float m[kMaxK], v[kMaxK], mo[kMaxK], vo[kMaxK];
#pragma unroll
  for (int i = 0; i < kMaxK; ++i) {
    if (i < k) {
      m[i] = d.mean[(size_t)b * k + i];
      v[i] = d.v[(size_t)b * k + i];
      mo[i] = d.old_mean[(size_t)b * k + i];
      vo[i] = d.old_v[(size_t)b * k + i];
    } else {
      m[i] = mo[i] = 0.f;
      v[i] = vo[i] = 1.f;
    }
  }
  // ---- mean part ---------------------------------------------------------------------------
  float maha = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxK; ++i)
    if (i < k) { const float t = (m[i] - mo[i]) / vo[i]; maha += t * t; }
  const float mean_part = d.proj_type == 0 ? 0.5f * maha : maha;
  float omega = 0.f;
  const bool mean_active = mean_part > d.eps_mean;
  if (mean_active) omega = fabsf(sqrtf(mean_part / d.eps_mean) - 1.0f);
#pragma unroll
  for (int i = 0; i < kMaxK; ++i)
    if (i < k) d.proj_mean[(size_t)b * k + i] = mean_active ? (m[i] + omega * mo[i]) / (1.0f + omega + 1e-16f) : m[i];

  d.eta[2 * (size_t)b + 1] = mean_active ? (double)omega : 0.0;  // branch flag for the backward pass

  // ---- covariance part -----------------------------------------------------------------------
  if (d.proj_type == 0) {
    double c[kMaxK], o[kMaxK];
#pragma unroll
    for (int i = 0; i < kMaxK; ++i) { c[i] = (double)(v[i] * v[i]); o[i] = (double)(vo[i] * vo[i]); }
    const double eps = (double)d.eps_cov;
    double eta = 0.0;
    KlEval f = kl_eval(c, o, k, 0.0);
    if (f.kl > eps) {
      double lo = 0.0, hi = 1.0;
      for (int it = 0; it < 200; ++it) {  // bracket: KL is decreasing in eta
        if (kl_eval(c, o, k, hi).kl <= eps) break;
        lo = hi;
        hi *= 2.0;
      }
      eta = 0.5 * (lo + hi);
      for (int it = 0; it < 100; ++it) {  // safeguarded Newton on KL(eta) - eps
        f = kl_eval(c, o, k, eta);
        const double res = f.kl - eps;
        if (res > 0.0) lo = eta; else hi = eta;
        if (fabs(res) <= 1e-14 * eps || (hi - lo) <= 1e-15 * hi) break;
        double nxt = eta - res / f.dkl;
        if (!(nxt > lo && nxt < hi)) nxt = 0.5 * (lo + hi);
        eta = nxt;
      }
    }
    d.eta[2 * (size_t)b] = eta;
#pragma unroll
    for (int i = 0; i < kMaxK; ++i) {
      if (i < k) {
        // reference: cov.new(float64 result) -> fp32, then .sqrt() in fp32 (kl_projection_layer.py:72)
        const float ct = eta > 0.0 ? (float)((eta + 1.0) / (eta / o[i] + 1.0 / c[i])) : v[i] * v[i];
        d.proj_v[(size_t)b * k + i] = sqrtf(ct);
      }
    }
  } else {
    float cov_part = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxK; ++i)
      if (i < k) { const float inv = 1.0f / vo[i]; cov_part += 1.0f + inv * (v[i] * v[i]) * inv - 2.0f * inv * v[i]; }
    float eta = 0.f;
    const bool active = cov_part > d.eps_cov;
    if (active) eta = fabsf(sqrtf(cov_part / d.eps_cov) - 1.0f);
    d.eta[2 * (size_t)b] = active ? (double)eta : 0.0;
#pragma unroll
    for (int i = 0; i < kMaxK; ++i)
      if (i < k) d.proj_v[(size_t)b * k + i] = active ? (v[i] + eta * vo[i]) / (1.0f + eta + 1e-16f) : v[i];
  }
}

__global__ void __launch_bounds__(128) trpl_bwd_kernel(const GrlProjDesc d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.batch) return;
  const int k = d.k;
  double m[kMaxK], v[kMaxK], mo[kMaxK], vo[kMaxK], gm[kMaxK], gv[kMaxK];
#pragma unroll
  for (int i = 0; i < kMaxK; ++i) {
    if (i < k) {
      m[i] = d.mean[(size_t)b * k + i];
      v[i] = d.v[(size_t)b * k + i];
      mo[i] = d.old_mean[(size_t)b * k + i];
      vo[i] = d.old_v[(size_t)b * k + i];
      gm[i] = d.grad_proj_mean[(size_t)b * k + i];
      gv[i] = d.grad_proj_v[(size_t)b * k + i];
    } else {
      m[i] = mo[i] = gm[i] = gv[i] = 0.0;
      v[i] = vo[i] = 1.0;
    }
  }
  // ---- mean: pm = (m + w mo) / (1 + w), w = sqrt(M / eps) - 1, M = scale * sum ((m - mo) / vo)^2
  const double scale = d.proj_type == 0 ? 0.5 : 1.0;
  double M = 0.0;
#pragma unroll
  for (int i = 0; i < kMaxK; ++i)
    if (i < k) { const double t = (m[i] - mo[i]) / vo[i]; M += t * t; }
  M *= scale;
  const double eps_m = (double)d.eps_mean;
  if (d.eta[2 * (size_t)b + 1] > 0.0) {
    const double w = sqrt(M / eps_m) - 1.0;
    double dot = 0.0;  // sum_i g_i d pm_i / d w
#pragma unroll
    for (int i = 0; i < kMaxK; ++i)
      if (i < k) dot += gm[i] * (mo[i] - m[i]) / ((1.0 + w) * (1.0 + w));
    const double dw_dM = 1.0 / (2.0 * eps_m * (1.0 + w));
#pragma unroll
    for (int i = 0; i < kMaxK; ++i)
      if (i < k)
        d.grad_mean[(size_t)b * k + i] =
            (float)(gm[i] / (1.0 + w) + dot * dw_dM * scale * 2.0 * (m[i] - mo[i]) / (vo[i] * vo[i]));
  } else {
#pragma unroll
    for (int i = 0; i < kMaxK; ++i)
      if (i < k) d.grad_mean[(size_t)b * k + i] = (float)gm[i];
  }
  // ---- covariance
  const double eta = d.eta[2 * (size_t)b];
  if (d.proj_type == 0) {
    // pv = sqrt(c~), c~ = (eta+1)/(eta/o + 1/c), c = v^2, o = vo^2; implicit gradient through eta(c)
    if (eta > 0.0) {
      double num = 0.0, den = 0.0;
      double gct[kMaxK], dct_dc[kMaxK], dkl_dct[kMaxK];
#pragma unroll
      for (int i = 0; i < kMaxK; ++i) {
        gct[i] = dct_dc[i] = dkl_dct[i] = 0.0;
        if (i < k) {
          const double c = v[i] * v[i], o = vo[i] * vo[i];
          const double D = eta / o + 1.0 / c;
          const double ct = (eta + 1.0) / D;
          gct[i] = gv[i] / (2.0 * sqrt(ct));
          dct_dc[i] = (eta + 1.0) / (D * D * c * c);
          const double dct_deta = (1.0 / c - 1.0 / o) / (D * D);
          dkl_dct[i] = 0.5 * (1.0 / o - 1.0 / ct);
          num += gct[i] * dct_deta;
          den += dkl_dct[i] * dct_deta;
        }
      }
#pragma unroll
      for (int i = 0; i < kMaxK; ++i)
        if (i < k) {
          const double gc = gct[i] * dct_dc[i] - num * (dkl_dct[i] * dct_dc[i]) / den;
          d.grad_v[(size_t)b * k + i] = (float)(gc * 2.0 * v[i]);
        }
    } else {
      // identity: pv = sqrt(v^2) = v
#pragma unroll
      for (int i = 0; i < kMaxK; ++i)
        if (i < k) d.grad_v[(size_t)b * k + i] = (float)gv[i];
    }
  } else {
    if (eta > 0.0) {
      // nv = (v + eta vo) / (1 + eta), eta = sqrt(S / eps) - 1, S = sum (v/vo - 1)^2
      double dot = 0.0;
#pragma unroll
      for (int i = 0; i < kMaxK; ++i)
        if (i < k) dot += gv[i] * (vo[i] - v[i]) / ((1.0 + eta) * (1.0 + eta));
      const double deta_dS = 1.0 / (2.0 * (double)d.eps_cov * (1.0 + eta));
#pragma unroll
      for (int i = 0; i < kMaxK; ++i)
        if (i < k)
          d.grad_v[(size_t)b * k + i] =
              (float)(gv[i] / (1.0 + eta) + dot * deta_dS * 2.0 * (v[i] / vo[i] - 1.0) / vo[i]);
    } else {
#pragma unroll
      for (int i = 0; i < kMaxK; ++i)
        if (i < k) d.grad_v[(size_t)b * k + i] = (float)gv[i];
    }
  }
}

}  // namespace grl

extern "C" {

static int check_proj(const GrlProjDesc* d, const char* who, bool bwd) {
  GRL_REQUIRE(d, GRL_EINVAL, "%s: null descriptor", who);
  GRL_REQUIRE(d->batch > 0 && d->k > 0, GRL_EINVAL, "%s: batch=%d k=%d", who, d->batch, d->k);
  GRL_REQUIRE(d->k <= GRL_MAX_ACTION_DIM, GRL_EUNSUPPORTED, "%s: k=%d > %d", who, d->k, GRL_MAX_ACTION_DIM);
  GRL_REQUIRE(d->proj_type == 0 || d->proj_type == 1, GRL_EUNSUPPORTED, "%s: proj_type=%d", who, d->proj_type);
  GRL_REQUIRE(d->eps_mean > 0.f && d->eps_cov > 0.f, GRL_EINVAL, "%s: bounds must be positive", who);
  GRL_REQUIRE(d->mean && d->v && d->old_mean && d->old_v && d->eta, GRL_EINVAL, "%s: null pointer", who);
  if (bwd) GRL_REQUIRE(d->grad_proj_mean && d->grad_proj_v && d->grad_mean && d->grad_v, GRL_EINVAL, "%s: null grad pointer", who);
  else GRL_REQUIRE(d->proj_mean && d->proj_v, GRL_EINVAL, "%s: null output pointer", who);
  return GRL_OK;
}

int grl_trpl_fwd(const GrlProjDesc* d, grl_stream_t stream) {
  const int rc = check_proj(d, "grl_trpl_fwd", false);
  if (rc != GRL_OK) return rc;
  grl::trpl_fwd_kernel<<<(d->batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_trpl_fwd");
}

int grl_trpl_bwd(const GrlProjDesc* d, grl_stream_t stream) {
  const int rc = check_proj(d, "grl_trpl_bwd", true);
  if (rc != GRL_OK) return rc;
  grl::trpl_bwd_kernel<<<(d->batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*d);
  return grl::check_launch("grl_trpl_bwd");
}

}  // extern "C"
