// Descendant-structured loss terms of calculate_loss that the reference's shipped scripts switch on
// (run_pipnet_20protos_multi_runs_seed42.sh: --tanh_desc "y|0.05", --minimize_contrasting_set 'y',
// --mask_prune_overspecific 'y|0|1.1'), batched over all nodes on the flat [V,P] pooled table:
//   tanh_desc     pipnet/train.py:1089-1133   tanh loss of every leaf below a node on the prototypes of the child it hangs under
//   contrast      pipnet/train.py:1017-1060   max activation of a child's prototypes over the node's OTHER descendants (TOPK 1)
//   mask pruning  pipnet/train.py:946-1015    overspecificity score x soft Gumbel presence mask + L1 of the mask
// The reference walks node -> child -> leaf in Python with a host sync per step; here every kernel is a flat grid over
// (leaf), (classifier weight element), (prototype) or (view, prototype) driven by the static tree tables.  All [V,P]-sized
// work: HBM/latency-trivial next to the GEMMs, written for zero host involvement (CUDA-graph capturable).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hc {

constexpr float REL_THRESH = 1e-3f;        // "relevant prototype" of a child: classifier weight > 1e-3  (:963, :1097)
constexpr float REL_THRESH_CS = 1e-5f;     // the contrasting-set term uses 1e-5                        (:1033)

struct DescWs {                // per-call workspace kept from forward to backward
  double* acc;                 // [7,N] td_sum, td_cnt, cs_sum, cs_cnt, ov_sum, l1_sum, rel_cnt
  int32_t* leader;             // [V]   first row with the same leaf label
  float* leafmax;              // [V,P] at leader rows: max over the leaf's rows
  int32_t* leafarg;            // [V,P] at leader rows: first row attaining it
};

struct DescParams {
  const float* pooled;         // [V,P]
  const float* wc;             // flat classifier weights
  const float* presence;       // [P,2] logits (mask pruning) or null
  const float* gumbel;         // [n_welems,2] Gumbel noise per (node, child, prototype)
  const long long* ys;         // [V] leaf index (sorted leaf-name order)
  const int8_t* tgt;           // [V,N] child label of the row's leaf at node n, -1 if not below n
  const int32_t* n_desc;       // [N]
  const int32_t *proto_off, *cls_off, *wc_off, *proto_node, *col_node, *welem_col, *welem_proto, *path_off, *path_col;
  int V, V_first, N, P, L, n_welems;
  int flags;
  float w_td, w_cs, w_ov, w_l1;     // already divided by N
  float eps, boost, inv_tau;        // boost <= 0: no boosting factor
};

constexpr int DESC_TANH = 1, DESC_CONTRAST = 2, DESC_MASK = 4, DESC_GEOMETRIC = 8, DESC_SG = 16;

__global__ void desc_leader_kernel(const long long* __restrict__ ys, int V, int32_t* __restrict__ leader) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < V; r += gridDim.x * blockDim.x) {
    const long long y = ys[r];
    int first = r;
    for (int q = 0; q < r; ++q)
      if (ys[q] == y) { first = q; break; }
    leader[r] = first;
  }
}

// leader rows only: max over the rows of the same leaf (both views), first row on ties
__global__ void desc_leafmax_kernel(const float* __restrict__ pooled, const long long* __restrict__ ys,
                                    const int32_t* __restrict__ leader, int V, int P, float* __restrict__ leafmax,
                                    int32_t* __restrict__ leafarg) {
  const int r = blockIdx.y;
  if (leader[r] != r) return;
  const long long y = ys[r];
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
    float m = pooled[(size_t)r * P + p];
    int a = r;
    for (int q = r + 1; q < V; ++q) {
      if (ys[q] != y) continue;
      const float v = pooled[(size_t)q * P + p];
      if (v > m) { m = v; a = q; }
    }
    leafmax[(size_t)r * P + p] = m;
    leafarg[(size_t)r * P + p] = a;
  }
}

// ------------------------------------------------------------------ tanh_desc forward: one block per leaf
__global__ void tanh_desc_fwd_kernel(DescParams q, double* __restrict__ acc) {
  const int d = blockIdx.x;
  __shared__ float red_a[2];
  __shared__ int red_c[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;      // 64 threads
  for (int i = q.path_off[d]; i < q.path_off[d + 1]; ++i) {
    const int col = q.path_col[i];
    const int n = q.col_node[col];
    if (q.n_desc[n] == 0) continue;                                // node skipped as a whole (:941-942)
    const int c = col - q.cls_off[n];
    const int p0 = q.proto_off[n], pn = q.proto_off[n + 1] - p0;
    const float* wrow = q.wc + q.wc_off[n] + (size_t)c * pn;
    float a = 0.f;
    int cnt = 0;
    for (int pl = threadIdx.x; pl < pn; pl += blockDim.x) {
      if (!(wrow[pl] > REL_THRESH)) continue;
      float s1 = 0.f, s2 = 0.f;
      for (int r = 0; r < q.V; ++r) {
        if (q.ys[r] != d) continue;
        const float v = q.pooled[(size_t)r * q.P + p0 + pl];
        if (r < q.V_first) s1 += v; else s2 += v;
      }
      a += logf(tanhf(s1) + q.eps) + logf(tanhf(s2) + q.eps);
      ++cnt;
    }
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) { red_a[warp] = a; red_c[warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
      const int ct = red_c[0] + red_c[1];
      if (ct > 0) {
        atomicAdd(acc + 0 * q.N + n, (double)(-0.5f * (red_a[0] + red_a[1]) / (float)ct));
        atomicAdd(acc + 1 * q.N + n, 1.0);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ contrasting set forward: thread per classifier element
__device__ __forceinline__ int contrast_argmax(const DescParams& q, int n, int c, int p, float* best_out) {
  float best = 0.f;
  int arg = -1;
  for (int r = 0; r < q.V; ++r) {
    const int tg = q.tgt[(size_t)r * q.N + n];
    if (tg < 0 || tg == c) continue;
    const float v = q.pooled[(size_t)r * q.P + p];
    if (arg < 0 || v > best) { best = v; arg = r; }
  }
  *best_out = best;
  return arg;
}

__global__ void contrast_fwd_kernel(DescParams q, double* __restrict__ acc) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= q.n_welems) return;
  const int col = q.welem_col[e];
  const int n = q.col_node[col];
  if (q.n_desc[n] == 0 || !(q.wc[e] > REL_THRESH_CS)) return;
  float best;
  const int arg = contrast_argmax(q, n, col - q.cls_off[n], q.welem_proto[e], &best);
  if (arg < 0) return;
  atomicAdd(acc + 2 * q.N + n, (double)best);
  atomicAdd(acc + 3 * q.N + n, 1.0);
}

// ------------------------------------------------------------------ mask pruning
// score of (node n, child c, prototype p): product over the leaves below c that occur in the batch of the leaf's max
// activation (boosted and clamped at 1, plain, or geometric mean).  Returns false when no such leaf is present.
struct ScoreInfo {
  float score;
  int n_leaves;
  int zeros;          // factors equal to 0 (for the "product of the others" in the backward)
  float nz_prod;      // product of the non-zero factors
};

__device__ __forceinline__ float mp_factor(const DescParams& q, float m, int n_leaves) {
  if (q.boost > 0.f) return fminf(m * q.boost, 1.0f);
  if (q.flags & DESC_GEOMETRIC) return powf(m, 1.0f / (float)n_leaves);
  return m;
}

__device__ __forceinline__ bool mp_score(const DescParams& q, const DescWs& ws, int n, int c, int p, ScoreInfo* out) {
  int nl = 0;
  for (int r = 0; r < q.V; ++r)
    if (ws.leader[r] == r && q.tgt[(size_t)r * q.N + n] == c) ++nl;
  out->n_leaves = nl;
  if (nl == 0) return false;
  float prod = 1.f, nz = 1.f;
  int zeros = 0;
  for (int r = 0; r < q.V; ++r) {
    if (ws.leader[r] != r || q.tgt[(size_t)r * q.N + n] != c) continue;
    const float f = mp_factor(q, ws.leafmax[(size_t)r * q.P + p], nl);
    prod *= f;
    if (f == 0.f) ++zeros; else nz *= f;
  }
  out->score = prod;
  out->zeros = zeros;
  out->nz_prod = nz;
  return true;
}

__device__ __forceinline__ bool child_present(const DescParams& q, int n, int c) {
  for (int r = 0; r < q.V; ++r)
    if (q.tgt[(size_t)r * q.N + n] == c) return true;
  return false;
}

// one step of the presence chain: y <- softmax((y + g) / tau) over the two entries (F.gumbel_softmax, hard=False)
__device__ __forceinline__ void gumbel_step(const DescParams& q, int e, float& y0, float& y1) {
  const float a0 = (y0 + q.gumbel[2 * (size_t)e]) * q.inv_tau;
  const float a1 = (y1 + q.gumbel[2 * (size_t)e + 1]) * q.inv_tau;
  const float m = fmaxf(a0, a1);
  const float e0 = __expf(a0 - m), e1 = __expf(a1 - m);
  const float inv = 1.0f / (e0 + e1);
  y0 = e0 * inv;
  y1 = e1 * inv;
}

__global__ void mask_prune_fwd_kernel(DescParams q, DescWs ws) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= q.P) return;
  const int n = q.proto_node[p];
  if (q.n_desc[n] == 0) return;
  const int p0 = q.proto_off[n], pn = q.proto_off[n + 1] - p0, pl = p - p0;
  const int cn = q.cls_off[n + 1] - q.cls_off[n];
  float y0 = q.presence[2 * (size_t)p], y1 = q.presence[2 * (size_t)p + 1];
  float ov = 0.f, l1 = 0.f;
  int rel_cnt = 0;
  for (int c = 0; c < cn; ++c) {
    const int e = q.wc_off[n] + c * pn + pl;
    const bool rel = q.wc[e] > REL_THRESH;
    rel_cnt += rel ? 1 : 0;                         // counted before the child is possibly skipped (:965)
    ScoreInfo si;
    if (!mp_score(q, ws, n, c, p, &si)) continue;   // no leaf of this child in the batch (:975-976)
    gumbel_step(q, e, y0, y1);                      // applied to whatever the previous child left (:978)
    if (rel) {
      ov -= si.score * y1;
      l1 += y1;
    }
  }
  if (ov != 0.f) atomicAdd(ws.acc + 4 * q.N + n, (double)ov);
  if (l1 != 0.f) atomicAdd(ws.acc + 5 * q.N + n, (double)l1);
  if (rel_cnt) atomicAdd(ws.acc + 6 * q.N + n, (double)rel_cnt);
}

// ------------------------------------------------------------------ combine: per-node statistics and the weighted sum
// stats[4,N]: tanh_desc (mean over leaves), contrast (mean over entries), overspecificity and mask-L1 (weighted, as the
// reference stores them :1006-1010); loss = sum_n w_td*td + w_cs*cs + ovsp + l1.
__global__ void desc_combine_kernel(DescParams q, const double* __restrict__ acc, float* __restrict__ stats,
                                    float* __restrict__ loss) {
  __shared__ double red[256];
  double part = 0.0;
  const int N = q.N;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float td = 0.f, cs = 0.f, ov = 0.f, l1 = 0.f;
    if (q.n_desc[n] > 0) {
      if ((q.flags & DESC_TANH) && acc[1 * N + n] > 0.0) td = (float)(acc[0 * N + n] / acc[1 * N + n]);
      if ((q.flags & DESC_CONTRAST) && acc[3 * N + n] > 0.0) cs = (float)(acc[2 * N + n] / acc[3 * N + n]);
      if ((q.flags & DESC_MASK) && acc[6 * N + n] > 0.0) {
        ov = q.w_ov * (float)(acc[4 * N + n] / acc[6 * N + n]);
        l1 = q.w_l1 * (float)(acc[5 * N + n] / acc[6 * N + n]);
      }
    }
    stats[0 * N + n] = td;
    stats[1 * N + n] = cs;
    stats[2 * N + n] = ov;
    stats[3 * N + n] = l1;
    part += (double)(q.w_td * td) + (double)(q.w_cs * cs) + (double)ov + (double)l1;
  }
  red[threadIdx.x] = part;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = (float)red[0];
}

// ------------------------------------------------------------------ backward w.r.t. pooled: thread per (row, prototype)
__global__ void desc_bwd_pooled_kernel(DescParams q, DescWs ws, const float* __restrict__ g_loss,
                                       float* __restrict__ g_pooled) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (p >= q.P) return;
  const int N = q.N;
  const int n = q.proto_node[p];
  const int tg = q.tgt[(size_t)r * N + n];
  float g = 0.f;
  if (tg >= 0 && q.n_desc[n] > 0) {
    const float up = g_loss[0];
    const int p0 = q.proto_off[n], pn = q.proto_off[n + 1] - p0, pl = p - p0;
    const int cn = q.cls_off[n + 1] - q.cls_off[n];
    const float* wn = q.wc + q.wc_off[n];
    const long long d = q.ys[r];
    // tanh_desc: the row feeds the term of its own leaf, through the prototypes of the child it hangs under
    if ((q.flags & DESC_TANH) && ws.acc[1 * N + n] > 0.0 && wn[(size_t)tg * pn + pl] > REL_THRESH) {
      const bool first = r < q.V_first;
      float s = 0.f;
      for (int k = first ? 0 : q.V_first; k < (first ? q.V_first : q.V); ++k)
        if (q.ys[k] == d) s += q.pooled[(size_t)k * q.P + p];
      int R = 0;
      for (int k = 0; k < pn; ++k) R += wn[(size_t)tg * pn + k] > REL_THRESH ? 1 : 0;
      const float t = tanhf(s);
      g += up * q.w_td / (float)ws.acc[1 * N + n] * (-0.5f / (float)R) * (1.f - t * t) / (t + q.eps);
    }
    // contrasting set: the row is a candidate for every OTHER child's prototypes
    if ((q.flags & DESC_CONTRAST) && ws.acc[3 * N + n] > 0.0) {
      for (int c = 0; c < cn; ++c) {
        if (c == tg || !(wn[(size_t)c * pn + pl] > REL_THRESH_CS)) continue;
        float best;
        if (contrast_argmax(q, n, c, p, &best) == r) g += up * q.w_cs / (float)ws.acc[3 * N + n];
      }
    }
    // mask pruning: the row carries the max of its leaf for this prototype
    if ((q.flags & DESC_MASK) && !(q.flags & DESC_SG) && ws.acc[6 * N + n] > 0.0 && wn[(size_t)tg * pn + pl] > REL_THRESH) {
      const int lead = ws.leader[r];
      if (ws.leafarg[(size_t)lead * q.P + p] == r) {
        ScoreInfo si;
        mp_score(q, ws, n, tg, p, &si);
        const float m = ws.leafmax[(size_t)lead * q.P + p];
        const float f = mp_factor(q, m, si.n_leaves);
        float dfac;                                   // d factor / d m
        if (q.boost > 0.f) dfac = (m * q.boost <= 1.0f) ? q.boost : 0.f;       // clamp(max=1) passes the gradient at <=
        else if (q.flags & DESC_GEOMETRIC) dfac = powf(m, 1.0f / (float)si.n_leaves - 1.0f) / (float)si.n_leaves;
        else dfac = 1.f;
        float others;                                 // product of the other leaves' factors
        if (f != 0.f) others = si.zeros ? 0.f : si.nz_prod / f;
        else others = si.zeros == 1 ? si.nz_prod : 0.f;
        // presence probability at this child's step of the chain
        float y0 = q.presence[2 * (size_t)p], y1 = q.presence[2 * (size_t)p + 1];
        for (int c = 0; c <= tg; ++c)
          if (child_present(q, n, c)) gumbel_step(q, q.wc_off[n] + c * pn + pl, y0, y1);
        g += up * q.w_ov / (float)ws.acc[6 * N + n] * (-y1) * others * dfac;
      }
    }
  }
  g_pooled[(size_t)r * q.P + p] = g;
}

// ------------------------------------------------------------------ backward w.r.t. the presence logits: thread per prototype
__global__ void mask_prune_bwd_presence_kernel(DescParams q, DescWs ws, const float* __restrict__ g_loss,
                                               float* __restrict__ g_presence) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= q.P) return;
  const int N = q.N;
  const int n = q.proto_node[p];
  float adj0 = 0.f, adj1 = 0.f;
  if (q.n_desc[n] > 0 && ws.acc[6 * N + n] > 0.0) {
    const float up = g_loss[0];
    const float inv_rel = 1.0f / (float)ws.acc[6 * N + n];
    const int p0 = q.proto_off[n], pn = q.proto_off[n + 1] - p0, pl = p - p0;
    const int cn = q.cls_off[n + 1] - q.cls_off[n];
    // reverse sweep over the applied children; the chain state at step c is recomputed from the start (C_n is small)
    for (int c = cn - 1; c >= 0; --c) {
      ScoreInfo si;
      if (!mp_score(q, ws, n, c, p, &si)) continue;
      float y0 = q.presence[2 * (size_t)p], y1 = q.presence[2 * (size_t)p + 1];
      for (int k = 0; k <= c; ++k)
        if (child_present(q, n, k)) gumbel_step(q, q.wc_off[n] + k * pn + pl, y0, y1);
      const int e = q.wc_off[n] + c * pn + pl;
      const float coef = (q.wc[e] > REL_THRESH) ? up * inv_rel * (q.w_l1 - q.w_ov * si.score) : 0.f;
      const float a1 = adj1 + coef, a0 = adj0;
      const float dlt = (a1 - a0) * y1 * (1.f - y1) * q.inv_tau;
      adj1 = dlt;
      adj0 = -dlt;
    }
  }
  g_presence[2 * (size_t)p] = adj0;
  g_presence[2 * (size_t)p + 1] = adj1;
}

}  // namespace hc
