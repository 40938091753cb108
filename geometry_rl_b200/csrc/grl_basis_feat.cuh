// Edge-basis invariant features for the tensor-core kernels (hepi.py:109-123 compute_invariants +
// ponita.py:233-244 PolynomialFeatures degree 2), software-pipelined over the persistent tile loop.
//
// Row r of a tile is (edge tile * 8 + (r >> 4), orientation r & 15).  Producing its 14 features needs two DEPENDENT
// global loads (edge -> node indices -> positions); done inside the tile they cost two exposed memory latencies per
// tile (ncu r01: the forward basis kernel was bound by exactly that, 3.3 us per tile).  Here thread r < 128 keeps
// the node indices of the tile after next and the positions of the next tile in registers, so every load is consumed
// one loop iteration after it was issued.
#pragma once
#include "grl_common.cuh"
#include "grl_tc.cuh"

namespace grl {

struct BasisFeatPipe {
  int es, ed;     // node indices of the tile after next (-1: no edge)
  float p[6];     // pos_src / pos_dst of the next tile
  bool valid;     // the next tile's row has an edge

  __device__ __forceinline__ void load_ids(const GrlBasisDesc& d, int tile) {
    const int r = threadIdx.x;
    const int e = tile * kTE + (r >> 4);
    es = -1; ed = -1;
    if (r < kTM && e < d.n_edges) { es = __ldg(d.edge_src + e); ed = __ldg(d.edge_dst + e); }
  }
  __device__ __forceinline__ void load_pos(const GrlBasisDesc& d) {
    valid = es >= 0;
    if (valid) {
      const float* ps = d.pos_src + 3 * (size_t)es;
      const float* pd = d.pos_dst + 3 * (size_t)ed;
      p[0] = __ldg(ps); p[1] = __ldg(ps + 1); p[2] = __ldg(ps + 2);
      p[3] = __ldg(pd); p[4] = __ldg(pd + 1); p[5] = __ldg(pd + 2);
    }
  }
  // prologue: positions of `tile`, indices of `tile + stride`
  __device__ __forceinline__ void start(const GrlBasisDesc& d, int tile, int stride) {
    load_ids(d, tile);
    load_pos(d);
    load_ids(d, tile + stride);
  }
  // features of the tile whose positions are in registers -> F image ([2 chunks][128 rows][8] bf16; features 14, 15 =
  // `one`, the bias columns), then advance: positions of tile + stride, indices of tile + 2 stride.
  __device__ __forceinline__ void emit_and_advance(const GrlBasisDesc& d, int tile, int stride, __nv_bfloat16* __restrict__ F,
                                                   float one) {
    const int r = threadIdx.x;
    if (r < kTM) {
      const int o = r & 15;
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = 0.f;
      if (valid) {
        const float rx = p[0] - p[3], ry = p[1] - p[4], rz = (d.dim == 3) ? p[2] - p[5] : 0.f;
        const float ox = d.ori[3 * o], oy = d.ori[3 * o + 1], oz = (d.dim == 3) ? d.ori[3 * o + 2] : 0.f;
        const float i1 = (rx * ox + ry * oy) + rz * oz;
        const float tx = rx - i1 * ox, ty = ry - i1 * oy, tz = rz - i1 * oz;
        const float i2 = sqrtf((tx * tx + ty * ty) + tz * tz);
        f[0] = i1; f[1] = i2;
        f[2] = i1 * i1; f[3] = i1 * i2; f[4] = i2 * i1; f[5] = i2 * i2;
        f[6] = f[2] * i1; f[7] = f[2] * i2; f[8] = f[3] * i1; f[9] = f[3] * i2;
        f[10] = f[4] * i1; f[11] = f[4] * i2; f[12] = f[5] * i1; f[13] = f[5] * i2;
        f[14] = one; f[15] = one;
      }
      *reinterpret_cast<uint4*>(F + ((size_t)0 * kTM + r) * 8) = tc::pack8(f);
      *reinterpret_cast<uint4*>(F + ((size_t)1 * kTM + r) * 8) = tc::pack8(f + 8);
    }
    load_pos(d);                       // uses the indices requested one iteration ago
    load_ids(d, tile + 2 * stride);
  }
};

}  // namespace grl
