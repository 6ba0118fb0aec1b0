// Row kernels for SPILL nodes: nodes whose raw logits the fused projection kernel only writes out (scratch matrix
// Zs[M, P_s], fp32) instead of finishing them in its epilogue.  They carry, for those nodes, everything the epilogue
// does for the fused ones -- softmax over the node's prototypes (pipnet/pipnet.py:146-147), global max-pool with
// first-occurrence argmax (:159, :24-32), the align loss (pipnet/train.py:1063-1069, :1399-1405) and, in the
// backward, dZ = S * (G - sum G*S) / tau -- reading Zs instead of TMEM; the backward needs no GEMM recompute.
//
//   narrow (P_n <= 64, "riders": nodes moved out of a nearly empty last tile into spare pad columns of the other
//           tiles, layout.py): one thread per location, the node's logits in registers -- the epilogue's own code
//           (softmax_row / pool_segment / add of the scattered pooled gradient) on data from global memory;
//   wide   (P_n > 64, e.g. 4 children x 20 prototypes, or flat trees with 20 x leaves prototypes): one warp per
//           location with the lanes striding over the prototypes (coalesced), per-row statistics (max, 1/sum, align
//           coefficient) kept in a small side table for the pooling pass and the backward.
#pragma once
#include "head_pair.cuh"

namespace hc {

// ------------------------------------------------------------------------------------------------ narrow nodes
// stand-alone launches of the rider row functions (head_pair.cuh) for the riders the fused kernels do not finish in
// their own tail (csrc/cabi.cu: run_pair folds the riders of the last segment class)
template <int S>
__global__ void __launch_bounds__(256) spill_narrow_fwd_kernel(const SpillParams p) {
  __shared__ uint4 xch_all[8][4 * PairCfg<S>::XQ];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((blockIdx.x * 8 + warp) * 32 >= p.halfM) return;                 // warp-uniform
  spill_narrow_fwd_rows<S>(p, (blockIdx.x * 8 + warp) * 32 + lane, lane, xch_all[warp]);
}
template <int S>
__global__ void __launch_bounds__(256) spill_narrow_bwd_kernel(const SpillParams p) {
  spill_narrow_bwd_row<S>(p, blockIdx.x * 256 + threadIdx.x);
}

// ------------------------------------------------------------------------------------------------ wide nodes
__device__ __forceinline__ float spill_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float spill_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// pass A: one warp per pair row (location of view 1 + the same location of view 2): row max / 1 over row sum of both
// views into `stats`, align term from the inner product of the two softmax rows.  grid: ceil(halfM / 8) blocks of 8 warps.
__global__ void __launch_bounds__(256) spill_wide_stats_kernel(const SpillParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_a = blockIdx.x * 8 + warp;
  if (row_a >= p.halfM) return;
  const bool valid_b = row_a < p.rowsB;
  const float* z1 = p.zs + (size_t)row_a * p.ldz + p.zoff;
  const float* z2 = p.zs + (size_t)(p.halfM + row_a) * p.ldz + p.zoff;
  float m1 = -INFINITY, m2 = -INFINITY;
  for (int c = lane; c < p.P_n; c += 32) {
    m1 = fmaxf(m1, __ldg(z1 + c));
    if (valid_b) m2 = fmaxf(m2, __ldg(z2 + c));
  }
  m1 = spill_warp_max(m1);
  m2 = spill_warp_max(m2);
  const float k1 = m1 * p.scale_log2, k2 = m2 * p.scale_log2;
  float l1 = 0.f, l2 = 0.f, e12 = 0.f;
  for (int c = lane; c < p.P_n; c += 32) {
    const float e1 = ex2(fmaf(__ldg(z1 + c), p.scale_log2, -k1));
    l1 += e1;
    if (valid_b) {
      const float e2 = ex2(fmaf(__ldg(z2 + c), p.scale_log2, -k2));
      l2 += e2;
      e12 = fmaf(e1, e2, e12);
    }
  }
  l1 = spill_warp_sum(l1);
  l2 = spill_warp_sum(l2);
  e12 = spill_warp_sum(e12);
  const float i1 = __frcp_rn(l1), i2 = valid_b ? __frcp_rn(l2) : 0.f;
  if (lane == 0) {
    p.stats[2 * (size_t)row_a] = m1;
    p.stats[2 * (size_t)row_a + 1] = i1;
    if (valid_b) {
      p.stats[2 * (size_t)(p.halfM + row_a)] = m2;
      p.stats[2 * (size_t)(p.halfM + row_a) + 1] = i2;
    }
    if (p.desc != nullptr && p.align_sum != nullptr && valid_b) {
      int v_a, loc;
      row_to_img(row_a, p.HW, p.inv_HW, v_a, loc);
      if (v_a < p.imgs_first && p.desc[(size_t)v_a * p.n_nodes + p.node]) {
        const float ip = e12 * i1 * i2;
        atomicAdd(p.align_sum + p.node, (double)(-__logf(ip + 1e-12f)));
      }
    }
  }
}

// pass B: max-pool.  One thread per (image, prototype of the node): walks the image's HW locations in order, so the
// first occurrence of the maximum wins without any tie-break arithmetic.  grid: (ceil(P_n / 128), V)
__global__ void __launch_bounds__(128) spill_wide_pool_kernel(const SpillParams p) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int v = blockIdx.y;
  if (c >= p.P_n) return;
  const size_t row0 = (size_t)v * p.HW;
  float best = -1.f;
  int arg = 0;
  for (int l = 0; l < p.HW; ++l) {
    const size_t r = row0 + l;
    const float m = p.stats[2 * r], inv = p.stats[2 * r + 1];
    const float s = ex2(fmaf(__ldg(p.zs + r * p.ldz + p.zoff + c), p.scale_log2, -m * p.scale_log2)) * inv;
    if (s > best) { best = s; arg = l; }
  }
  p.pooled_packed[(size_t)v * p.P + p.poff + c] =
      ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)arg);
}

// backward: one warp per pair row.  G = -ca * S_other + [location is the argmax] * g_pooled;  dZ = S * (G - <G, S>) / tau
__global__ void __launch_bounds__(256) spill_wide_bwd_kernel(const SpillParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_a = blockIdx.x * 8 + warp;
  if (row_a >= p.halfM) return;
  const bool valid_b = row_a < p.rowsB;
  const size_t r1 = row_a, r2 = (size_t)p.halfM + row_a;
  const float* z1 = p.zs + r1 * p.ldz + p.zoff;
  const float* z2 = p.zs + r2 * p.ldz + p.zoff;
  int v_a, loc;
  row_to_img(row_a, p.HW, p.inv_HW, v_a, loc);
  const float m1 = p.stats[2 * r1], i1 = p.stats[2 * r1 + 1];
  const float m2 = valid_b ? p.stats[2 * r2] : 0.f, i2 = valid_b ? p.stats[2 * r2 + 1] : 0.f;
  const float k1 = m1 * p.scale_log2, k2 = m2 * p.scale_log2;
  const int2* sc1 = p.scat + (size_t)v_a * p.P + p.poff;
  const int2* sc2 = p.scat + (size_t)(v_a + p.imgs_first) * p.P + p.poff;
  // inner product of the two rows (align coefficient) and the two <G, S> dots
  float ip = 0.f;
  if (valid_b)
    for (int c = lane; c < p.P_n; c += 32)
      ip = fmaf(ex2(fmaf(__ldg(z1 + c), p.scale_log2, -k1)) * i1, ex2(fmaf(__ldg(z2 + c), p.scale_log2, -k2)) * i2, ip);
  ip = spill_warp_sum(ip);
  float ca = 0.f;
  if (valid_b && v_a < p.imgs_first && p.coef_align != nullptr) ca = p.coef_align[(size_t)v_a * p.n_nodes + p.node] * __frcp_rn(ip + 1e-12f);
  float d1 = 0.f, d2 = 0.f;
  for (int c = lane; c < p.P_n; c += 32) {
    const float s1 = ex2(fmaf(__ldg(z1 + c), p.scale_log2, -k1)) * i1;
    const float s2 = valid_b ? ex2(fmaf(__ldg(z2 + c), p.scale_log2, -k2)) * i2 : 0.f;
    const int2 e1 = __ldg(sc1 + c);
    float g1 = -ca * s2 + (e1.x == loc ? __int_as_float(e1.y) : 0.f);
    d1 = fmaf(g1, s1, d1);
    if (valid_b) {
      const int2 e2 = __ldg(sc2 + c);
      float g2 = -ca * s1 + (e2.x == loc ? __int_as_float(e2.y) : 0.f);
      d2 = fmaf(g2, s2, d2);
    }
  }
  d1 = spill_warp_sum(d1);
  d2 = spill_warp_sum(d2);
  __nv_bfloat16* o1 = p.dz + r1 * p.P_c + p.dz_col;
  __nv_bfloat16* o2 = p.dz + r2 * p.P_c + p.dz_col;
  for (int c = lane; c < p.dz_width; c += 32) {
    float a = 0.f, b = 0.f;
    if (c < p.P_n) {
      const float s1 = ex2(fmaf(__ldg(z1 + c), p.scale_log2, -k1)) * i1;
      const float s2 = valid_b ? ex2(fmaf(__ldg(z2 + c), p.scale_log2, -k2)) * i2 : 0.f;
      const int2 e1 = __ldg(sc1 + c);
      const float g1 = -ca * s2 + (e1.x == loc ? __int_as_float(e1.y) : 0.f);
      a = s1 * (g1 - d1) * p.inv_tau;
      if (valid_b) {
        const int2 e2 = __ldg(sc2 + c);
        const float g2 = -ca * s1 + (e2.x == loc ? __int_as_float(e2.y) : 0.f);
        b = s2 * (g2 - d2) * p.inv_tau;
      }
    }
    o1[c] = __float2bfloat16(a);
    if (valid_b) o2[c] = __float2bfloat16(b);
  }
}

}  // namespace hc
