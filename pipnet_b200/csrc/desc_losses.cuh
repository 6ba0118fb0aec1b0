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
  double* acc;                 // [5,N] td_sum, cs_sum, cs_cnt, ov_sum, l1_sum
  int32_t* leader;             // [V]   first row with the same leaf label
  int32_t* next;               // [V]   next row with the same leaf label, -1 at the end of the list
  float* leaf_s1;              // [V,P] at leader rows: sum over the leaf's rows of the first view half
  float* leaf_s2;              // [V,P] ... of the second half
  float* leaf_max;             // [V,P] at leader rows: max over the leaf's rows (both halves)
  int32_t* leaf_arg;           // [V,P] at leader rows: first row attaining it
  int32_t* col_rel;            // [K]   prototypes of the child column with classifier weight > 1e-3
  int32_t* col_present;        // [K]   distinct leaves of the batch below the child
  int32_t* cs_arg;             // [E]   contrasting set: argmax row of classifier element e, -1 if none
  float* score;                // [E]   mask pruning: overspecificity score of (node, child, prototype)
  float* nz_prod;              // [E]   product of its non-zero factors
  int32_t* zeros;              // [E]   number of zero factors
  float* y1_at;                // [E]   presence probability right after this child's Gumbel step
};

struct DescParams {
  const float* pooled;         // [V,P]
  const float* wc;             // flat classifier weights [E]
  const float* presence;       // [P,2] logits (mask pruning) or null
  const float* gumbel;         // [E,2] Gumbel noise per (node, child, prototype)
  const long long* ys;         // [V] leaf index (sorted leaf-name order)
  const int8_t* tgt;           // [V,N] child label of the row's leaf at node n, -1 if not below n
  const int32_t* n_desc;       // [N]
  const int32_t *proto_off, *cls_off, *wc_off, *proto_node, *col_node, *welem_col, *welem_proto, *col_nleaves;
  int V, V_first, N, P, K, E;
  int flags;
  float w_td, w_cs, w_ov, w_l1;     // already divided by N
  float eps, boost, inv_tau;        // boost <= 0: no boosting factor
};

constexpr int DESC_TANH = 1, DESC_CONTRAST = 2, DESC_MASK = 4, DESC_GEOMETRIC = 8, DESC_SG = 16;
constexpr int ACC_TD = 0, ACC_CS = 1, ACC_CSN = 2, ACC_OV = 3, ACC_L1 = 4;

// ------------------------------------------------------------------ batch structure: rows grouped by leaf
// One block; ys staged in shared memory.  leader[r] = first row with the label of r, next[r] = next such row.
__global__ void desc_prep_kernel(const long long* __restrict__ ys, int V, int32_t* __restrict__ leader,
                                 int32_t* __restrict__ next) {
  extern __shared__ long long sy[];
  for (int r = threadIdx.x; r < V; r += blockDim.x) sy[r] = ys[r];
  __syncthreads();
  for (int r = threadIdx.x; r < V; r += blockDim.x) {
    const long long y = sy[r];
    int first = r, nx = -1;
    for (int q = 0; q < r; ++q)
      if (sy[q] == y) { first = q; break; }
    for (int q = r + 1; q < V; ++q)
      if (sy[q] == y) { nx = q; break; }
    leader[r] = first;
    next[r] = nx;
  }
}

// leader rows only: per-half sums, max and first argmax over the rows of the leaf (list walk, ~2-4 rows)
__global__ void desc_leaf_stats_kernel(DescParams q, DescWs ws) {
  const int r = blockIdx.y;
  if (ws.leader[r] != r) return;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= q.P) return;
  float s1 = 0.f, s2 = 0.f, m = -1.f;
  int a = r;
  for (int k = r; k >= 0; k = ws.next[k]) {
    const float v = q.pooled[(size_t)k * q.P + p];
    if (k < q.V_first) s1 += v; else s2 += v;
    if (v > m) { m = v; a = k; }
  }
  const size_t o = (size_t)r * q.P + p;
  ws.leaf_s1[o] = s1;
  ws.leaf_s2[o] = s2;
  ws.leaf_max[o] = m;
  ws.leaf_arg[o] = a;
}

// one warp per child column: relevant prototype count and number of distinct batch leaves below the child
__global__ void desc_col_stats_kernel(DescParams q, DescWs ws) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= q.K) return;
  const int n = q.col_node[k], c = k - q.cls_off[n];
  const int pn = q.proto_off[n + 1] - q.proto_off[n];
  const float* wrow = q.wc + q.wc_off[n] + (size_t)c * pn;
  int rel = 0, pres = 0;
  for (int pl = lane; pl < pn; pl += 32) rel += wrow[pl] > REL_THRESH ? 1 : 0;
  for (int r = lane; r < q.V; r += 32) pres += (ws.leader[r] == r && q.tgt[(size_t)r * q.N + n] == c) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) {
    rel += __shfl_xor_sync(0xffffffffu, rel, o);
    pres += __shfl_xor_sync(0xffffffffu, pres, o);
  }
  if (lane == 0) { ws.col_rel[k] = rel; ws.col_present[k] = pres; }
}

__device__ __forceinline__ float mp_factor(const DescParams& q, float m, int n_leaves) {
  if (q.boost > 0.f) return fminf(m * q.boost, 1.0f);
  if (q.flags & DESC_GEOMETRIC) return powf(m, 1.0f / (float)n_leaves);
  return m;
}

// one warp per classifier weight element e = (node n, child c, prototype p), lanes over the rows of the batch:
//   contrasting set : max / first argmax of pooled[r,p] over rows below n but NOT below c      (pipnet/train.py:1044-1050)
//   mask pruning    : product over the batch leaves below c of factor(max over the leaf's rows)  (:968-985)
__global__ void desc_elem_reduce_kernel(DescParams q, DescWs ws) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= q.E) return;
  const int col = q.welem_col[e], p = q.welem_proto[e];
  const int n = q.col_node[col], c = col - q.cls_off[n];
  const bool live = q.n_desc[n] > 0;
  const bool do_cs = (q.flags & DESC_CONTRAST) && live && q.wc[e] > REL_THRESH_CS;
  const bool do_mp = (q.flags & DESC_MASK) && live;
  const int nl = do_mp ? ws.col_present[col] : 0;
  float best = -1.f, prod = 1.f, nz = 1.f;
  int arg = 0x7fffffff, zeros = 0;
  if (do_cs || (do_mp && nl > 0)) {
    for (int r = lane; r < q.V; r += 32) {
      const int tg = q.tgt[(size_t)r * q.N + n];
      if (do_cs && tg >= 0 && tg != c) {
        const float v = q.pooled[(size_t)r * q.P + p];
        if (v > best) { best = v; arg = r; }
      }
      if (do_mp && tg == c && ws.leader[r] == r) {
        const float f = mp_factor(q, ws.leaf_max[(size_t)r * q.P + p], nl);
        prod *= f;
        if (f == 0.f) ++zeros; else nz *= f;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
      prod *= __shfl_xor_sync(0xffffffffu, prod, o);
      nz *= __shfl_xor_sync(0xffffffffu, nz, o);
      zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    }
  }
  if (lane == 0) {
    const bool found = do_cs && arg != 0x7fffffff;
    if (q.flags & DESC_CONTRAST) ws.cs_arg[e] = found ? arg : -1;
    if (found) {
      atomicAdd(ws.acc + ACC_CS * q.N + n, (double)best);
      atomicAdd(ws.acc + ACC_CSN * q.N + n, 1.0);
    }
    if (q.flags & DESC_MASK) { ws.score[e] = prod; ws.nz_prod[e] = nz; ws.zeros[e] = zeros; }
  }
}

// ------------------------------------------------------------------ tanh_desc forward: thread per (batch leaf, node)
// term(n, leaf) = -1/2 * sum_halves mean_{p in R(child)} log(tanh(sum of the leaf's rows) + eps); leaves that are not in
// the batch contribute the constant -log(eps) each and are added in closed form by the combine kernel.
__global__ void tanh_desc_fwd_kernel(DescParams q, DescWs ws) {
  const int u = blockIdx.y;
  if (ws.leader[u] != u) return;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= q.N || q.n_desc[n] == 0) return;
  const int c = q.tgt[(size_t)u * q.N + n];
  if (c < 0) return;
  const int rel = ws.col_rel[q.cls_off[n] + c];
  if (rel == 0) return;
  const int p0 = q.proto_off[n], pn = q.proto_off[n + 1] - p0;
  const float* wrow = q.wc + q.wc_off[n] + (size_t)c * pn;
  const float* s1 = ws.leaf_s1 + (size_t)u * q.P + p0;
  const float* s2 = ws.leaf_s2 + (size_t)u * q.P + p0;
  float a = 0.f;
  for (int pl = 0; pl < pn; ++pl)
    if (wrow[pl] > REL_THRESH) a += logf(tanhf(s1[pl]) + q.eps) + logf(tanhf(s2[pl]) + q.eps);
  atomicAdd(ws.acc + ACC_TD * q.N + n, (double)(-0.5f * a / (float)rel));
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

// thread per prototype: the Gumbel chain over the node's children, overspecificity and mask-L1 sums
__global__ void mask_prune_fwd_kernel(DescParams q, DescWs ws) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= q.P) return;
  const int n = q.proto_node[p];
  if (q.n_desc[n] == 0) return;
  const int p0 = q.proto_off[n], pn = q.proto_off[n + 1] - p0, pl = p - p0;
  const int k0 = q.cls_off[n], cn = q.cls_off[n + 1] - k0;
  float y0 = q.presence[2 * (size_t)p], y1 = q.presence[2 * (size_t)p + 1];
  float ov = 0.f, l1 = 0.f;
  for (int c = 0; c < cn; ++c) {
    if (ws.col_present[k0 + c] == 0) continue;      // no leaf of this child in the batch (:975-976)
    const int e = q.wc_off[n] + c * pn + pl;
    gumbel_step(q, e, y0, y1);                      // applied to whatever the previous child left (:978)
    ws.y1_at[e] = y1;
    if (q.wc[e] > REL_THRESH) {
      ov -= ws.score[e] * y1;
      l1 += y1;
    }
  }
  if (ov != 0.f) atomicAdd(ws.acc + ACC_OV * q.N + n, (double)ov);
  if (l1 != 0.f) atomicAdd(ws.acc + ACC_L1 * q.N + n, (double)l1);
}

// per-node denominators shared by forward and backward
__device__ __forceinline__ void node_counts(const DescParams& q, const DescWs& ws, int n, int* td_cnt, int* td_absent,
                                            int* rel_total) {
  int cnt = 0, absent = 0, rel = 0;
  for (int k = q.cls_off[n]; k < q.cls_off[n + 1]; ++k) {
    const int r = ws.col_rel[k];
    rel += r;                                       // counted whether or not the child is in the batch (:965)
    if (r > 0) {
      cnt += q.col_nleaves[k];
      absent += q.col_nleaves[k] - ws.col_present[k];
    }
  }
  *td_cnt = cnt;
  *td_absent = absent;
  *rel_total = rel;
}

// ------------------------------------------------------------------ combine: per-node statistics and the weighted sum
// stats[4,N]: tanh_desc (mean over leaves), contrast (mean over entries), overspecificity and mask-L1 (weighted, as the
// reference stores them :1006-1010); loss = sum_n w_td*td + w_cs*cs + ovsp + l1.
__global__ void desc_combine_kernel(DescParams q, DescWs ws, float* __restrict__ stats, float* __restrict__ loss) {
  __shared__ double red[256];
  double part = 0.0;
  const int N = q.N;
  const double* acc = ws.acc;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float td = 0.f, cs = 0.f, ov = 0.f, l1 = 0.f;
    if (q.n_desc[n] > 0) {
      int td_cnt, td_absent, rel_total;
      node_counts(q, ws, n, &td_cnt, &td_absent, &rel_total);
      if ((q.flags & DESC_TANH) && td_cnt > 0)
        td = (float)((acc[ACC_TD * N + n] + (double)td_absent * (double)(-logf(q.eps))) / (double)td_cnt);
      if ((q.flags & DESC_CONTRAST) && acc[ACC_CSN * N + n] > 0.0) cs = (float)(acc[ACC_CS * N + n] / acc[ACC_CSN * N + n]);
      if ((q.flags & DESC_MASK) && rel_total > 0) {
        ov = q.w_ov * (float)(acc[ACC_OV * N + n] / (double)rel_total);
        l1 = q.w_l1 * (float)(acc[ACC_L1 * N + n] / (double)rel_total);
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
    const int k0 = q.cls_off[n], cn = q.cls_off[n + 1] - k0;
    const int e_own = q.wc_off[n] + tg * pn + pl;
    const bool rel_own = q.wc[e_own] > REL_THRESH;
    const int lead = ws.leader[r];
    int td_cnt = 0, td_absent = 0, rel_total = 0;
    if (q.flags & (DESC_TANH | DESC_MASK)) node_counts(q, ws, n, &td_cnt, &td_absent, &rel_total);
    // tanh_desc: the row feeds the term of its own leaf, through the prototypes of the child it hangs under
    if ((q.flags & DESC_TANH) && rel_own && td_cnt > 0) {
      const float s = (r < q.V_first ? ws.leaf_s1 : ws.leaf_s2)[(size_t)lead * q.P + p];
      const float t = tanhf(s);
      g += up * q.w_td / (float)td_cnt * (-0.5f / (float)ws.col_rel[k0 + tg]) * (1.f - t * t) / (t + q.eps);
    }
    // contrasting set: the row is a candidate for every OTHER child's prototypes
    if ((q.flags & DESC_CONTRAST) && ws.acc[ACC_CSN * N + n] > 0.0) {
      int hits = 0;
      for (int c = 0; c < cn; ++c)
        if (c != tg && ws.cs_arg[q.wc_off[n] + c * pn + pl] == r) ++hits;
      if (hits) g += (float)hits * up * q.w_cs / (float)ws.acc[ACC_CSN * N + n];
    }
    // mask pruning: the row carries the max of its leaf for this prototype
    if ((q.flags & DESC_MASK) && !(q.flags & DESC_SG) && rel_own && rel_total > 0 &&
        ws.leaf_arg[(size_t)lead * q.P + p] == r) {
      const int nl = ws.col_present[k0 + tg];
      const float m = ws.leaf_max[(size_t)lead * q.P + p];
      const float f = mp_factor(q, m, nl);
      float dfac;                                   // d factor / d m
      if (q.boost > 0.f) dfac = (m * q.boost <= 1.0f) ? q.boost : 0.f;        // clamp(max=1) passes the gradient at <=
      else if (q.flags & DESC_GEOMETRIC) dfac = powf(m, 1.0f / (float)nl - 1.0f) / (float)nl;
      else dfac = 1.f;
      const int zeros = ws.zeros[e_own];
      float others;                                 // product of the other leaves' factors
      if (f != 0.f) others = zeros ? 0.f : ws.nz_prod[e_own] / f;
      else others = zeros == 1 ? ws.nz_prod[e_own] : 0.f;
      g += up * q.w_ov / (float)rel_total * (-ws.y1_at[e_own]) * others * dfac;
    }
  }
  g_pooled[(size_t)r * q.P + p] = g;
}

// ------------------------------------------------------------------ backward w.r.t. the presence logits: thread per prototype
__global__ void mask_prune_bwd_presence_kernel(DescParams q, DescWs ws, const float* __restrict__ g_loss,
                                               float* __restrict__ g_presence) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= q.P) return;
  const int n = q.proto_node[p];
  float adj0 = 0.f, adj1 = 0.f;
  if (q.n_desc[n] > 0) {
    int td_cnt, td_absent, rel_total;
    node_counts(q, ws, n, &td_cnt, &td_absent, &rel_total);
    if (rel_total > 0) {
      const float up = g_loss[0];
      const float inv_rel = 1.0f / (float)rel_total;
      const int p0 = q.proto_off[n], pn = q.proto_off[n + 1] - p0, pl = p - p0;
      const int k0 = q.cls_off[n], cn = q.cls_off[n + 1] - k0;
      for (int c = cn - 1; c >= 0; --c) {           // reverse sweep over the applied Gumbel steps
        if (ws.col_present[k0 + c] == 0) continue;
        const int e = q.wc_off[n] + c * pn + pl;
        const float y1 = ws.y1_at[e];
        const float coef = (q.wc[e] > REL_THRESH) ? up * inv_rel * (q.w_l1 - q.w_ov * ws.score[e]) : 0.f;
        const float dlt = ((adj1 + coef) - adj0) * y1 * (1.f - y1) * q.inv_tau;
        adj1 = dlt;
        adj0 = -dlt;
      }
    }
  }
  g_presence[2 * (size_t)p] = adj0;
  g_presence[2 * (size_t)p + 1] = adj1;
}

}  // namespace hc
