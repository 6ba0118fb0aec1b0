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

struct SpillParams {
  const float* zs;          // [M, ldz]
  int ldz;
  int M, halfM, rowsB, HW, P, n_nodes, imgs_first;
  float scale_log2, inv_tau, inv_HW;
  // node record
  int node, P_n, poff, zoff, dz_col, dz_width;
  // forward
  unsigned long long* pooled_packed;
  double* align_sum;
  const uint8_t* desc;
  // backward
  const int2* scat;
  const float* coef_align;
  __nv_bfloat16* dz;
  int P_c;
  float* stats;             // wide nodes: [M, 2] {row max, 1 / row sum}; forward writes, pooling + backward read
};

__device__ __forceinline__ void row_to_img(int row, int HW, float inv_HW, int& v, int& loc) {
  v = __float2int_rz(__int2float_rz(row) * inv_HW);
  loc = row - v * HW;
  while (loc < 0) { --v; loc += HW; }
  while (loc >= HW) { ++v; loc -= HW; }
}

// ------------------------------------------------------------------------------------------------ narrow, forward
// grid: ceil(halfM / 32) warps (8 warps per block); lane = one location of view 1 and the same location of view 2
template <int S>
__global__ void __launch_bounds__(256) spill_narrow_fwd_kernel(const SpillParams p) {
  __shared__ uint4 xch_all[8][2 * PairCfg<S>::XQ];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_a = (blockIdx.x * 8 + warp) * 32 + lane;
  if ((blockIdx.x * 8 + warp) * 32 >= p.halfM) return;                 // warp-uniform
  const bool valid_a = row_a < p.halfM, valid_b = row_a < p.rowsB;
  int v_a, loc;
  row_to_img(row_a, p.HW, p.inv_HW, v_a, loc);
  const int v_first = __shfl_sync(0xffffffffu, v_a, 0);
  const int loc_first = __shfl_sync(0xffffffffu, loc, 0);
  const bool has_boundary = __ballot_sync(0xffffffffu, v_a != v_first) != 0u;
  const int lane_b = p.HW - loc_first;
  const int nv_a = __popc(__ballot_sync(0xffffffffu, valid_a));
  const int nv_b = __popc(__ballot_sync(0xffffffffu, valid_b));
  uint32_t ra[S], rb[S];
#pragma unroll
  for (int i = 0; i < S; ++i) ra[i] = rb[i] = 0u;
  const int n4 = (p.P_n + 3) >> 2;
  if (valid_a) {
    const float4* z = reinterpret_cast<const float4*>(p.zs + (size_t)row_a * p.ldz + p.zoff);
#pragma unroll
    for (int i = 0; i < S / 4; ++i)
      if (i < n4) { const float4 t = __ldg(z + i); ra[4*i] = __float_as_uint(t.x); ra[4*i+1] = __float_as_uint(t.y); ra[4*i+2] = __float_as_uint(t.z); ra[4*i+3] = __float_as_uint(t.w); }
  }
  if (valid_b) {
    const float4* z = reinterpret_cast<const float4*>(p.zs + (size_t)(p.halfM + row_a) * p.ldz + p.zoff);
#pragma unroll
    for (int i = 0; i < S / 4; ++i)
      if (i < n4) { const float4 t = __ldg(z + i); rb[4*i] = __float_as_uint(t.x); rb[4*i+1] = __float_as_uint(t.y); rb[4*i+2] = __float_as_uint(t.z); rb[4*i+3] = __float_as_uint(t.w); }
  }
  float s1[S], s2[S];
  softmax_row<S, true>(ra, p.P_n, p.scale_log2, s1);
  softmax_row<S, true>(rb, p.P_n, p.scale_log2, s2);
  float ip = 0.f;
  {
    float ip4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < S; i += 2)
      fma2(ip4[i & 3], ip4[(i & 3) + 1], s1[i], s1[i + 1], s2[i], s2[i + 1], ip4[i & 3], ip4[(i & 3) + 1]);
    ip = (ip4[0] + ip4[1]) + (ip4[2] + ip4[3]);
  }
  if (p.desc != nullptr && p.align_sum != nullptr) {
    float a = 0.f;
    if (valid_a && valid_b && v_a < p.imgs_first && p.desc[(size_t)v_a * p.n_nodes + p.node]) a = -__logf(ip + 1e-12f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0 && a != 0.f) atomicAdd(p.align_sum + p.node, (double)a);
  }
  uint4* xch = xch_all[warp];
  unsigned long long* t1 = p.pooled_packed + (size_t)v_first * p.P + p.poff;
  unsigned long long* t2 = t1 + (size_t)p.imgs_first * p.P;
  if (nv_a == 32 && !has_boundary) pool_segment_fast<S>(s1, loc_first, p.P_n, lane, xch, t1);
  else if (nv_a > 0) pool_segment<S>(s1, valid_a, v_a, v_first, has_boundary, loc_first, lane_b, p.P_n, lane, t1, p.P);
  if (nv_b == 32 && !has_boundary) pool_segment_fast<S>(s2, loc_first, p.P_n, lane, xch, t2);
  else if (nv_b > 0) pool_segment<S>(s2, valid_b, v_a, v_first, has_boundary, loc_first, lane_b, p.P_n, lane, t2, p.P);
}

// ------------------------------------------------------------------------------------------------ narrow, backward
template <int S>
__device__ __forceinline__ void narrow_dz_row(const SpillParams& p, const float* s, const float* s_other, float ca, int v_img,
                                              int loc, int row) {
  float g[S];
  float dot = 0.f;
  const int2* sc = p.scat + (size_t)v_img * p.P + p.poff;
#pragma unroll
  for (int i = 0; i < S; ++i) {
    g[i] = -ca * s_other[i];
    if (i < p.P_n) {
      const int2 e = __ldg(sc + i);          // {argmax location of (image, prototype), pooled gradient}
      if (e.x == loc) g[i] += __int_as_float(e.y);
    }
    dot = fmaf(g[i], s[i], dot);
  }
  __nv_bfloat16* out = p.dz + (size_t)row * p.P_c + p.dz_col;
  // masked columns (k >= P_n) have s[k] = 0, so their dZ is an exact zero; columns past S (the 8-column rounding of the
  // node's dZ block, or the pitch padding it owns) are zero-filled explicitly
#pragma unroll
  for (int k = 0; k < S; k += 2)
    if (k < p.dz_width)
      *reinterpret_cast<uint32_t*>(out + k) = pack_bf16x2(s[k] * (g[k] - dot) * p.inv_tau, s[k + 1] * (g[k + 1] - dot) * p.inv_tau);
  for (int k = S; k < p.dz_width; k += 2) *reinterpret_cast<uint32_t*>(out + k) = 0u;
}

template <int S>
__global__ void __launch_bounds__(256) spill_narrow_bwd_kernel(const SpillParams p) {
  const int row_a = blockIdx.x * 256 + threadIdx.x;
  if (row_a >= p.halfM) return;
  const bool valid_b = row_a < p.rowsB;
  int v_a, loc;
  row_to_img(row_a, p.HW, p.inv_HW, v_a, loc);
  uint32_t ra[S], rb[S];
#pragma unroll
  for (int i = 0; i < S; ++i) ra[i] = rb[i] = 0u;
  const int n4 = (p.P_n + 3) >> 2;
  {
    const float4* z = reinterpret_cast<const float4*>(p.zs + (size_t)row_a * p.ldz + p.zoff);
#pragma unroll
    for (int i = 0; i < S / 4; ++i)
      if (i < n4) { const float4 t = __ldg(z + i); ra[4*i] = __float_as_uint(t.x); ra[4*i+1] = __float_as_uint(t.y); ra[4*i+2] = __float_as_uint(t.z); ra[4*i+3] = __float_as_uint(t.w); }
  }
  if (valid_b) {
    const float4* z = reinterpret_cast<const float4*>(p.zs + (size_t)(p.halfM + row_a) * p.ldz + p.zoff);
#pragma unroll
    for (int i = 0; i < S / 4; ++i)
      if (i < n4) { const float4 t = __ldg(z + i); rb[4*i] = __float_as_uint(t.x); rb[4*i+1] = __float_as_uint(t.y); rb[4*i+2] = __float_as_uint(t.z); rb[4*i+3] = __float_as_uint(t.w); }
  }
  float s1[S], s2[S];
  softmax_row<S, true>(ra, p.P_n, p.scale_log2, s1);
  softmax_row<S, true>(rb, p.P_n, p.scale_log2, s2);
  float ip = 0.f;
  {
    float ip4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < S; i += 2)
      fma2(ip4[i & 3], ip4[(i & 3) + 1], s1[i], s1[i + 1], s2[i], s2[i + 1], ip4[i & 3], ip4[(i & 3) + 1]);
    ip = (ip4[0] + ip4[1]) + (ip4[2] + ip4[3]);
  }
  float ca = 0.f;
  if (valid_b && v_a < p.imgs_first && p.coef_align != nullptr) ca = p.coef_align[(size_t)v_a * p.n_nodes + p.node] * __frcp_rn(ip + 1e-12f);
  narrow_dz_row<S>(p, s1, s2, ca, v_a, loc, row_a);
  if (valid_b) narrow_dz_row<S>(p, s2, s1, ca, v_a + p.imgs_first, loc, p.halfM + row_a);
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
