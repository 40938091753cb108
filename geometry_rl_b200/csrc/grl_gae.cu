// K3: generalized advantage estimation as a reverse-time warp scan.
//   delta_t = r_t + gamma (1 - terminated_t) V_{t+1} - V_t
//   A_t     = delta_t + gamma lambda (1 - done_t) A_{t+1}
//   target  = A_t + V_t
// Replaces the arithmetic of torchrl GAE(shifted=True) at examples/torchrl/train.py:134-140,249-252.
// One warp per environment; lane l of chunk k owns time t = T-1-32k-l, so loads are coalesced and the
// first-order recurrence becomes an inclusive scan over affine maps A -> b + a A (Kogge-Stone, 5 steps).
#include "grl_common.cuh"

namespace grl {

__global__ void __launch_bounds__(256) gae_scan_kernel(const float* __restrict__ reward, const float* __restrict__ value,
                                                      const uint8_t* __restrict__ done, const uint8_t* __restrict__ term,
                                                      float gamma, float lmbda, int B, int T, float* __restrict__ adv,
                                                      float* __restrict__ vtarget) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float gl = gamma * lmbda;
  for (int b = blockIdx.x * warps_per_block + (threadIdx.x >> 5); b < B; b += gridDim.x * warps_per_block) {
    const float* r = reward + (size_t)b * T;
    const float* v = value + (size_t)b * (T + 1);
    const uint8_t* dn = done + (size_t)b * T;
    const uint8_t* tm = term + (size_t)b * T;
    float carry = 0.f;  // A_{t+1} entering the chunk
    for (int hi = T - 1; hi >= 0; hi -= 32) {
      const int t = hi - lane;
      float a = 1.f, bb = 0.f, vt = 0.f;
      if (t >= 0) {
        vt = v[t];
        const float not_term = tm[t] ? 0.f : 1.f;
        const float not_done = dn[t] ? 0.f : 1.f;
        bb = r[t] + gamma * not_term * v[t + 1] - vt;
        a = gl * not_done;
      }
      // inclusive scan of affine maps: lane l ends with f_l o f_{l-1} o ... o f_0
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float ap = __shfl_up_sync(0xffffffffu, a, d);
        const float bp = __shfl_up_sync(0xffffffffu, bb, d);
        if (lane >= d) {
          bb = fmaf(a, bp, bb);
          a = a * ap;
        }
      }
      const float A = fmaf(a, carry, bb);
      if (t >= 0) {
        adv[(size_t)b * T + t] = A;
        vtarget[(size_t)b * T + t] = A + vt;
      }
      carry = __shfl_sync(0xffffffffu, A, 31);  // earliest time of this chunk (identity lanes pass A through)
    }
  }
}

}  // namespace grl

extern "C" int grl_gae_scan(const float* reward, const float* value_T1, const uint8_t* done, const uint8_t* terminated,
                            float gamma, float lmbda, int B, int T, float* advantage, float* value_target,
                            grl_stream_t stream) {
  GRL_REQUIRE(reward && value_T1 && done && terminated && advantage && value_target, GRL_EINVAL,
              "grl_gae_scan: null pointer");
  GRL_REQUIRE(B > 0 && T > 0, GRL_EINVAL, "grl_gae_scan: B=%d T=%d", B, T);
  const int warps_per_block = 8;
  int grid = (B + warps_per_block - 1) / warps_per_block;
  const int cap = 16 * grl::sm_count();
  if (grid > cap) grid = cap;
  grl::gae_scan_kernel<<<grid, 32 * warps_per_block, 0, (cudaStream_t)stream>>>(reward, value_T1, done, terminated, gamma,
                                                                                lmbda, B, T, advantage, value_target);
  return grl::check_launch("grl_gae_scan");
}
