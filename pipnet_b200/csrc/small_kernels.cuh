// Small SIMT kernels around the fused projection: weight packing, pool unpacking, the per-node
// non-negative classifier, class / tanh / orthogonality losses (forward + backward) and the joint
// leaf distribution.  All operate on FLAT tensors ([V,P] prototypes, [V,K] child logits with
// K = sum of children over nodes) driven by small int32 node tables, replacing the reference's
// per-node Python loops (pipnet/pipnet.py:124-170, pipnet/train.py:933-1194).
#pragma once
#include "ptx.cuh"
#include "head_pair.cuh"   // rider row code (spill_narrow_fwd_rows) used by the forward finish kernel

namespace hc {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float block_sum(float v, float* sh /* >= 32 floats */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) sh[0] = r;
  __syncthreads();
  r = sh[0];
  return r;
}

// ---------------------------------------------------------------- packing / casting
// Wp[r, :] = bf16(W[row_map[r], :]) or 0 for padding rows.  8 channels per thread.
__global__ void pack_weights_kernel(const float* __restrict__ w, const int32_t* __restrict__ row_map, int P_pad, int C,
                                    __nv_bfloat16* __restrict__ wp) {
  const int c8 = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)P_pad * c8) return;
  const int r = int(idx / c8), c = int(idx - (long long)r * c8) * 8;
  const int src = row_map[r];
  uint4 o = make_uint4(0, 0, 0, 0);
  if (src >= 0) {
    const float4 a = *reinterpret_cast<const float4*>(w + (size_t)src * C + c);
    const float4 b = *reinterpret_cast<const float4*>(w + (size_t)src * C + c + 4);
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
  }
  *reinterpret_cast<uint4*>(wp + (size_t)r * C + c) = o;
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i];
    const float4 b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
}

// fp32 -> three bf16 terms hi + mid + lo (24 significant bits), written as stacked planes dst[t*n + i].
// Used by the fp32-accurate projection: X*W ~ sum of six bf16 cross products accumulated in fp32.
__device__ __forceinline__ void split3(float x, __nv_bfloat16& a, __nv_bfloat16& b, __nv_bfloat16& c) {
  a = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(a);
  b = __float2bfloat16_rn(r1);
  c = __float2bfloat16_rn(r1 - __bfloat162float(b));
}
__global__ void split3_f32_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    __nv_bfloat16 a, b, c;
    split3(src[i], a, b, c);
    dst[i] = a; dst[n + i] = b; dst[2 * n + i] = c;
  }
}
// same for the prototype kernels, combined with the tile padding: Wp3[t*P_pad + r, :] = term t of W[row_map[r], :]
__global__ void pack_weights_split3_kernel(const float* __restrict__ w, const int32_t* __restrict__ row_map, int P_pad,
                                           int C, __nv_bfloat16* __restrict__ wp3) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)P_pad * C) return;
  const int r = int(idx / C), c = int(idx - (long long)r * C);
  const int src = row_map[r];
  __nv_bfloat16 a = __float2bfloat16(0.f), b = a, d = a;
  if (src >= 0) split3(w[(size_t)src * C + c], a, b, d);
  const size_t plane = (size_t)P_pad * C;
  wp3[idx] = a; wp3[plane + idx] = b; wp3[2 * plane + idx] = d;
}

// NCHW fp32/bf16 -> NHWC bf16 rows (ResNet features are NCHW-contiguous).  32x32 smem transpose.
template <typename T>
__global__ void nchw_to_rows_bf16_kernel(const T* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C, int HW) {
  __shared__ float tile[32][33];
  const int v = blockIdx.z;
  const int c0 = blockIdx.y * 32, l0 = blockIdx.x * 32;
  const T* s = src + (size_t)v * C * HW;
  __nv_bfloat16* d = dst + (size_t)v * HW * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, l = l0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && l < HW) ? float(s[(size_t)c * HW + l]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int l = l0 + i, c = c0 + threadIdx.x;
    if (c < C && l < HW) d[(size_t)l * C + c] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

// ---------------------------------------------------------------- backbone hand-off
// X rows (bf16, channels-last) = layer_scale[c] * keep[v] * y[row, c] + residual[row, c]: the tail of the LAST ConvNeXt
// block (torchvision CNBlock.forward: layer_scale * block(input), stochastic depth, += input; the reference uses the
// stock model, features/convnext_features.py:18-25, util/args.py:503 names features.7.2).  One pass writes the feature
// matrix K1's TMA reads -- no separate scale / add / cast / layout kernels between the backbone and the head.
// y, residual: channels-last rows [M, C], fp32 or bf16; 8 channels (16 output bytes) per thread.
template <typename TY, typename TR>
__global__ void __launch_bounds__(256) scale_residual_rows_kernel(const TY* __restrict__ y, const TR* __restrict__ res,
                                                                  const float* __restrict__ gamma, const float* __restrict__ keep,
                                                                  long long n8, int C8, int HW, __nv_bfloat16* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / C8;
    const int c = int(i - row * C8) * 8;
    const float k = keep != nullptr ? keep[row / HW] : 1.f;
    float yv[8], rv[8];
    if constexpr (sizeof(TY) == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(y) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(y) + 2 * i + 1);
      yv[0] = a.x; yv[1] = a.y; yv[2] = a.z; yv[3] = a.w; yv[4] = b.x; yv[5] = b.y; yv[6] = b.z; yv[7] = b.w;
    } else {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(y) + i);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); yv[2 * j] = f.x; yv[2 * j + 1] = f.y; }
    }
    if constexpr (sizeof(TR) == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(res) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(res) + 2 * i + 1);
      rv[0] = a.x; rv[1] = a.y; rv[2] = a.z; rv[3] = a.w; rv[4] = b.x; rv[5] = b.y; rv[6] = b.z; rv[7] = b.w;
    } else {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(res) + i);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); rv[2 * j] = f.x; rv[2 * j + 1] = f.y; }
    }
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaf(g[2 * j] * k, yv[2 * j], rv[2 * j]), fmaf(g[2 * j + 1] * k, yv[2 * j + 1], rv[2 * j + 1]));
      ow[j] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    reinterpret_cast<uint4*>(out)[i] = o;
  }
}

// ---------------------------------------------------------------- pool unpack
// packed = (float bits << 32) | (0xFFFFFFFF - location)  ->  pooled, argmax (+ inference threshold,
// pipnet/pipnet.py:168-169: pooled < 0.1 -> 0).
__global__ void unpack_pool_kernel(const unsigned long long* __restrict__ packed, long long n, float thresh,
                                   float* __restrict__ pooled, int32_t* __restrict__ argmax) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = packed[i];
  float v = __uint_as_float((uint32_t)(k >> 32));
  if (v < thresh) v = 0.f;
  pooled[i] = v;
  argmax[i] = (int32_t)(0xFFFFFFFFu - (uint32_t)k);
}

// ---------------------------------------------------------------- label tables
// tgt[v,n] = child label of sample v at node n, or -1 (pipnet/train.py:934-937) from the static
// leaf->node table anc[L,N]; n_desc[n] = number of samples below node n; desc = tgt >= 0 for the
// first-half images.
__global__ void label_tables_kernel(const long long* __restrict__ ys, const int8_t* __restrict__ anc, int V, int V_first,
                                    int N, int L, int8_t* __restrict__ tgt, uint8_t* __restrict__ desc,
                                    int32_t* __restrict__ n_desc /* zeroed by the caller */) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= V * N) return;
  const int v = idx / N, n = idx - v * N;
  const long long y = ys[v];
  const int8_t t = (y >= 0 && y < L) ? anc[(size_t)y * N + n] : int8_t(-1);
  tgt[idx] = t;
  if (v < V_first) desc[idx] = (t >= 0);
  if (t >= 0) atomicAdd(n_desc + n, 1);
}

// ---------------------------------------------------------------- classifier (NonNegLinear, pipnet/pipnet.py:1035-1036)
// out[v, k] = sum_p relu(Wc[k-th row]) * pooled[v, proto_off[node]+p] (+ bias[k]); one thread per (v, k).
// col_node[k] = node of flat child column k; Wc rows are stored row-major per node at wc_off[node].
__global__ void classifier_fwd_kernel(const float* __restrict__ pooled, const float* __restrict__ wc,
                                      const float* __restrict__ bias, const int32_t* __restrict__ col_node,
                                      const int32_t* __restrict__ proto_off, const int32_t* __restrict__ cls_off,
                                      const int32_t* __restrict__ wc_off, int V, int P, int K, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= V * K) return;
  const int v = idx / K, k = idx - v * K;
  const int n = col_node[k];
  const int p0 = proto_off[n], pn = proto_off[n + 1] - p0;
  const float* w = wc + wc_off[n] + (size_t)(k - cls_off[n]) * pn;
  const float* x = pooled + (size_t)v * P + p0;
  float acc = 0.f;
  for (int p = 0; p < pn; ++p) acc = fmaf(fmaxf(w[p], 0.f), x[p], acc);
  out[idx] = acc + (bias ? bias[k] : 0.f);
}

// g_pooled[v, p] += sum_c g_out[v, c] * relu(Wc[c, p]); one thread per (v, p).
__global__ void classifier_bwd_pooled_kernel(const float* __restrict__ g_out, const float* __restrict__ wc,
                                             const int32_t* __restrict__ proto_node, const int32_t* __restrict__ proto_off,
                                             const int32_t* __restrict__ cls_off, const int32_t* __restrict__ wc_off, int V,
                                             int P, int K, float* __restrict__ g_pooled, int accumulate) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)V * P) return;
  const int v = int(idx / P), p = int(idx - (long long)v * P);
  const int n = proto_node[p];
  const int p0 = proto_off[n], pn = proto_off[n + 1] - p0;
  const int k0 = cls_off[n], kn = cls_off[n + 1] - k0;
  float acc = 0.f;
  for (int c = 0; c < kn; ++c) acc = fmaf(g_out[(size_t)v * K + k0 + c], fmaxf(wc[wc_off[n] + (size_t)c * pn + (p - p0)], 0.f), acc);
  g_pooled[idx] = accumulate ? g_pooled[idx] + acc : acc;
}

// g_Wc[c, p] = [Wc > 0] * sum_v g_out[v, c] * pooled[v, p]; one warp per weight element.
__global__ void classifier_bwd_weight_kernel(const float* __restrict__ g_out, const float* __restrict__ pooled,
                                             const float* __restrict__ wc, const int32_t* __restrict__ welem_col,
                                             const int32_t* __restrict__ welem_proto, int V, int P, int K, int n_w,
                                             float* __restrict__ g_wc, float* __restrict__ g_bias) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= n_w) return;
  const int k = welem_col[wid], p = welem_proto[wid];
  float acc = 0.f;
  if (wc[wid] > 0.f)
    for (int v = lane; v < V; v += 32) acc = fmaf(g_out[(size_t)v * K + k], pooled[(size_t)v * P + p], acc);
  acc = warp_sum(acc);
  if (lane == 0) g_wc[wid] = acc;
  (void)g_bias;
}
__global__ void classifier_bwd_bias_kernel(const float* __restrict__ g_out, int V, int K, float* __restrict__ g_bias) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float acc = 0.f;
  for (int v = 0; v < V; ++v) acc += g_out[(size_t)v * K + k];
  g_bias[k] = acc;
}

// ---------------------------------------------------------------- class loss (pipnet/train.py:1153-1163, util/custom_losses.py:22-34)
// x = log1p(out ** mult) with mult = net._multiplier (pipnet/train.py:1158; main_dist.py:426 initialises it to 2 and
// freezes it -- that value takes the exact out*out path), or x = out when mult == 0 (pipnet_sparsity n).
__device__ __forceinline__ float sparsity_x(float o, float mult) {
  return mult == 2.f ? log1pf(o * o) : (mult > 0.f ? log1pf(powf(o, mult)) : o);
}
__device__ __forceinline__ float sparsity_dx(float o, float mult) {
  if (mult == 2.f) return 2.f * o / (1.f + o * o);
  return mult > 0.f ? mult * powf(o, mult - 1.f) / (1.f + powf(o, mult)) : 1.f;
}
// One block per node.
// loss[n] = mean over descendants of w[t] * (logsumexp(x) - x[t]); also per-node accuracy counts.
__global__ void class_loss_fwd_kernel(const float* __restrict__ out, const int8_t* __restrict__ tgt,
                                      const float* __restrict__ child_w, const int32_t* __restrict__ cls_off,
                                      const int32_t* __restrict__ n_desc, int V, int N, int K, float mult,
                                      float* __restrict__ loss, int32_t* __restrict__ n_correct) {
  __shared__ float sh[32];
  const int n = blockIdx.x;
  const int k0 = cls_off[n], kn = cls_off[n + 1] - k0;
  float acc = 0.f, corr = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const int t = tgt[(size_t)v * N + n];
    if (t < 0) continue;
    const float* o = out + (size_t)v * K + k0;
    float mx = -INFINITY, best = -INFINITY;
    int arg = 0;
    for (int c = 0; c < kn; ++c) {
      const float x = sparsity_x(o[c], mult);
      mx = fmaxf(mx, x);
      if (o[c] > best) { best = o[c]; arg = c; }     // torch.max(node_logits, 1): first max (pipnet/train.py:1189)
    }
    float se = 0.f, xt = 0.f;
    for (int c = 0; c < kn; ++c) {
      const float x = sparsity_x(o[c], mult);
      se += expf(x - mx);
      if (c == t) xt = x;
    }
    acc += child_w[k0 + t] * (logf(se) + mx - xt);
    corr += (arg == t) ? 1.f : 0.f;
  }
  acc = block_sum(acc, sh);
  corr = block_sum(corr, sh);
  if (threadIdx.x == 0) {
    const int nd = n_desc[n];
    loss[n] = nd > 0 ? acc / float(nd) : 0.f;
    n_correct[n] = int(corr + 0.5f);
  }
}
// g_out[v, k] = g_loss[n] / n_desc * w[t] * (softmax(x)[c] - [c == t]) * dx/dout; one thread per (v, k).
__global__ void class_loss_bwd_kernel(const float* __restrict__ out, const int8_t* __restrict__ tgt,
                                      const float* __restrict__ child_w, const int32_t* __restrict__ col_node,
                                      const int32_t* __restrict__ cls_off, const int32_t* __restrict__ n_desc,
                                      const float* __restrict__ g_loss, int V, int N, int K, float mult,
                                      float* __restrict__ g_out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= V * K) return;
  const int v = idx / K, k = idx - v * K;
  const int n = col_node[k];
  const int t = tgt[(size_t)v * N + n];
  float g = 0.f;
  if (t >= 0) {
    const int k0 = cls_off[n], kn = cls_off[n + 1] - k0;
    const float* o = out + (size_t)v * K + k0;
    float mx = -INFINITY;
    for (int c = 0; c < kn; ++c) mx = fmaxf(mx, sparsity_x(o[c], mult));
    float se = 0.f;
    for (int c = 0; c < kn; ++c) se += expf(sparsity_x(o[c], mult) - mx);
    const float ok = o[k - k0];
    const float xk = sparsity_x(ok, mult);
    const float sm = expf(xk - mx) / se;
    const float dx = sparsity_dx(ok, mult);
    g = g_loss[n] / float(n_desc[n]) * child_w[k0 + t] * (sm - ((k - k0) == t ? 1.f : 0.f)) * dx;
  }
  g_out[idx] = g;
}

// ---------------------------------------------------------------- tanh loss (pipnet/train.py:1076-1087)
// Per node: -1/2 * sum_{view half h} mean_p log(tanh(sum_{desc b in h} pooled[b,p]) + eps).  One block per node.
// colsum[h*P + p] keeps the masked column sums for the backward.
// grid (N, 2 view halves), 256 threads: warps stride over the half's rows, lanes over the node's prototypes.
// part[h*N + n] = -1/2 * mean_p log(tanh(colsum)+eps)  (0 for nodes without descendants, pipnet/train.py:941-942)
__global__ void tanh_loss_fwd_kernel(const float* __restrict__ pooled, const int8_t* __restrict__ tgt,
                                     const int32_t* __restrict__ proto_off, const int32_t* __restrict__ n_desc, int V,
                                     int V_first, int N, int P, float eps, float* __restrict__ part,
                                     float* __restrict__ colsum) {
  __shared__ float sh[8][64];
  __shared__ float red[32];
  const int n = blockIdx.x, h = blockIdx.y;
  const int p0 = proto_off[n], pn = proto_off[n + 1] - p0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vb = h ? V_first : 0, ve = h ? V : V_first;
  float acc = 0.f;
  for (int pc = 0; pc < pn; pc += 64) {          // 64 prototypes per pass (two per lane)
    float t0 = 0.f, t1 = 0.f;
    const int pa = pc + lane, pb = pc + 32 + lane;
    for (int v = vb + warp; v < ve; v += 8) {
      if (tgt[(size_t)v * N + n] >= 0) {
        const float* row = pooled + (size_t)v * P + p0;
        if (pa < pn) t0 += row[pa];
        if (pb < pn) t1 += row[pb];
      }
    }
    sh[warp][lane] = t0;
    sh[warp][32 + lane] = t1;
    __syncthreads();
    if (threadIdx.x < 64 && pc + threadIdx.x < pn) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
      colsum[(size_t)h * P + p0 + pc + threadIdx.x] = t;
      acc += logf(tanhf(t) + eps);
    }
    __syncthreads();
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) part[h * N + n] = n_desc[n] > 0 ? -0.5f * acc / float(pn) : 0.f;
}
// g_pooled[v,p] (+)= g_loss[n] * (-1/(2 P_n)) * (1 - th^2) / (th + eps) for descendant rows; thread per (v,p).
__global__ void tanh_loss_bwd_kernel(const float* __restrict__ colsum, const int8_t* __restrict__ tgt,
                                     const int32_t* __restrict__ proto_node, const int32_t* __restrict__ proto_off,
                                     const float* __restrict__ g_loss, int V, int V_first, int N, int P, float eps,
                                     float* __restrict__ g_pooled, int accumulate) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)V * P) return;
  const int v = int(idx / P), p = int(idx - (long long)v * P);
  const int n = proto_node[p];
  float g = 0.f;
  if (tgt[(size_t)v * N + n] >= 0) {
    const int pn = proto_off[n + 1] - proto_off[n];
    const float th = tanhf(colsum[(size_t)(v >= V_first) * P + p]);
    g = g_loss[n] * (-0.5f / float(pn)) * (1.f - th * th) / (th + eps);
  }
  g_pooled[idx] = accumulate ? g_pooled[idx] + g : g;
}

// ---------------------------------------------------------------- kernel-orthogonality loss (pipnet/train.py:1136-1151, orth_dist :1408-1412)
// Per node: rows of W whose classifier column has any weight > 1e-3; E = W_rel W_rel^T - I (P_rel < C);
// loss = ||E||_F.  The Gram matrix is tiny (P_n^2 dots of length C) but latency-critical inside the step, so it is
// spread over the whole GPU: ONE WARP per entry (n, i, j >= i) of the [N, P_max, P_max] slab, operands read straight from
// L2 (all prototype kernels are 1.5 MB) with independent coalesced loads; the slab is kept for the backward.
// A block-per-node version that staged W_n in shared memory spent 10 of its 17 us in the serial copy loop.
__global__ void __launch_bounds__(256, 5) orth_gram_kernel(const float* __restrict__ w, const float* __restrict__ wc,
                                 const int32_t* __restrict__ proto_off, const int32_t* __restrict__ cls_off,
                                 const int32_t* __restrict__ wc_off, int N, int C, int P_max, float* __restrict__ E,
                                 uint8_t* __restrict__ rel) {
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int slots = P_max * P_max;
  const int tri = P_max * (P_max + 1) / 2;            // E is symmetric: warps only for the upper triangle (j >= i),
  const int n = int(wid / tri);                       // mirrored below -- half the warps, one wave on cub27
  if (n >= N) return;
  int rem = int(wid - (long long)n * tri), i = 0;
  while (rem >= P_max - i) { rem -= P_max - i; ++i; }
  const int j = i + rem;
  const int p0 = proto_off[n], pn = proto_off[n + 1] - p0;
  if (i >= pn || j >= pn) return;
  const int kn = cls_off[n + 1] - cls_off[n];
  const float* wcn = wc + wc_off[n];
  bool ri = false, rj = false;                        // relevant = some classifier weight > 1e-3 (pipnet/train.py:1140)
  for (int c = 0; c < kn; ++c) {
    ri |= wcn[(size_t)c * pn + i] > 0.001f;
    rj |= wcn[(size_t)c * pn + j] > 0.001f;
  }
  if (i == j && lane == 0) rel[p0 + i] = ri;
  // the dot product does not wait for the relevance flags (its loads are issued beside theirs; the kernel is load latency)
  const float* a = w + (size_t)(p0 + i) * C;
  const float* b = w + (size_t)(p0 + j) * C;
  float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int c = lane;
  for (; c + 224 < C; c += 256) {
#pragma unroll
    for (int u = 0; u < 8; ++u) d[u] = fmaf(a[c + 32 * u], b[c + 32 * u], d[u]);
  }
  for (; c < C; c += 32) d[0] = fmaf(a[c], b[c], d[0]);
  const float dot = warp_sum(((d[0] + d[1]) + (d[2] + d[3])) + ((d[4] + d[5]) + (d[6] + d[7])));
  const float e = (ri && rj) ? dot - (i == j ? 1.f : 0.f) : 0.f;
  if (lane == 0) {
    float* En = E + (size_t)n * slots;
    En[i * P_max + j] = e;
    En[j * P_max + i] = e;
  }
}
// sumsq[n] = ||E_n||_F^2 in a fixed order (one warp per node): deterministic, unlike per-entry atomics
__global__ void orth_sumsq_kernel(const float* __restrict__ E, const int32_t* __restrict__ proto_off, int N, int P_max,
                                  float* __restrict__ sumsq) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const int pn = proto_off[n + 1] - proto_off[n];
  const float* En = E + (size_t)n * P_max * P_max;
  float acc = 0.f;
  for (int ij = lane; ij < pn * pn; ij += 32) {
    const float e = En[(ij / pn) * P_max + (ij % pn)];
    acc = fmaf(e, e, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) sumsq[n] = acc;
}
// dW[i,:] = g[n] * (2/L) * sum_j E[i,j] W[j,:]   (E symmetric, zero outside the relevant rows): one thread per
// (prototype row, channel); writes every element (zeros where the term is off), so the caller does not clear g_w.
__global__ void orth_bwd_kernel(const float* __restrict__ w, const int32_t* __restrict__ proto_node,
                                const int32_t* __restrict__ proto_off, int C, int P_max, const float* __restrict__ loss,
                                const float* __restrict__ E, const uint8_t* __restrict__ rel,
                                const float* __restrict__ g_loss, float* __restrict__ g_w) {
  const int row = blockIdx.x;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int n = proto_node[row];
  const int p0 = proto_off[n], pn = proto_off[n + 1] - p0;
  const float L = loss[n], g = g_loss[n];
  float acc = 0.f;
  if (rel[row] && L > 0.f && g != 0.f) {
    const float* Er = E + ((size_t)n * P_max + (row - p0)) * P_max;
    const float* wn = w + (size_t)p0 * C + c;
#pragma unroll 4
    for (int j = 0; j < pn; ++j) acc = fmaf(Er[j], wn[(size_t)j * C], acc);
    acc *= g * 2.f / L;
  }
  g_w[(size_t)row * C + c] = acc;
}

// ---------------------------------------------------------------- loss combination (one block)
// stats[0..3][n] = per-node align / tanh / orth / class loss (0 for nodes without descendants or disabled terms);
// total = sum_n sum_k weight[k] * stats[k][n]  (weights already contain the 1/N of pipnet/train.py:1071 etc.)
struct LossWeights { float w[4]; };
__global__ void loss_combine_kernel(const float* __restrict__ align, const float* __restrict__ tanh_part,
                                    const float* __restrict__ orth_sq, const float* __restrict__ cls,
                                    const int32_t* __restrict__ n_desc, int N, LossWeights lw, float* __restrict__ stats,
                                    float* __restrict__ total) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const bool on = n_desc[n] > 0;
    const float a = (align && on) ? align[n] : 0.f;
    const float t = (tanh_part && on) ? tanh_part[n] + tanh_part[N + n] : 0.f;
    const float o = (orth_sq && on) ? sqrtf(orth_sq[n]) : 0.f;
    const float c = (cls && on) ? cls[n] : 0.f;
    stats[n] = a; stats[N + n] = t; stats[2 * N + n] = o; stats[3 * N + n] = c;
    acc += lw.w[0] * a + lw.w[1] * t + lw.w[2] * o + lw.w[3] * c;
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) *total = acc;
}
// gvec[k][n] = g_total * weight[k] (upstream gradient of every per-node term)
__global__ void loss_grads_kernel(const float* __restrict__ g_total, int N, LossWeights lw, float* __restrict__ gvec) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 4 * N) return;
  gvec[idx] = g_total[0] * lw.w[idx / N];
}

// ---------------------------------------------------------------- align loss finalize / backward prep
// loss[n] = align_sum[n] / ((n_desc/2) * HW)   (mean over rows of the masked view-1 images, pipnet/train.py:1403)
__global__ void align_finalize_kernel(const double* __restrict__ align_sum, const int32_t* __restrict__ n_desc, int N, int HW,
                                      float* __restrict__ loss) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int nd = n_desc[n] / 2;
  loss[n] = nd > 0 ? float(align_sum[n] / (double(nd) * double(HW))) : 0.f;
}
// coef[b,n] = desc[b,n] * g_align[n] * 0.5 / ((n_desc/2) * HW): each side of the symmetric loss gets half
// (the other side is detached, pipnet/train.py:1068-1069).
__global__ void align_coef_kernel(const uint8_t* __restrict__ desc, const int32_t* __restrict__ n_desc,
                                  const float* __restrict__ g_align, int B, int N, int HW, float* __restrict__ coef) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N) return;
  const int n = idx % N;
  const int nd = n_desc[n] / 2;
  coef[idx] = (desc[idx] && nd > 0) ? g_align[n] * 0.5f / (float(nd) * float(HW)) : 0.f;
}
__global__ void make_scat_kernel(const int32_t* __restrict__ argmax, const float* __restrict__ g_pooled,
                                 const float* __restrict__ pooled, float thresh, long long n, int2* __restrict__ scat) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g = g_pooled[i];
  if (pooled != nullptr && pooled[i] < thresh) g = 0.f;   // inference threshold kills the gradient too
  scat[i] = make_int2(argmax[i], __float_as_int(g));
}

// ================================================================ fused chains (ABI v6)
// The per-step small work between the four large kernels used to be ~27 launches on the critical path (36 us of an
// 0.31 ms cub27 step, almost all of it launch latency).  The same arithmetic in FOUR multi-role launches: a block's role
// is a range of blockIdx.x, roles never communicate inside a launch (except the last-block-done combine of the losses).

// ---- block-activity tables of dZ (block-sparse backward GEMMs) -------------------------------------------------------
// With hierarchical labels an image drives only the nodes on its root-to-leaf path, so most (image, node) blocks of
// dZ[M, P_c] are exactly zero (68 % on cub27, 95 % on cub190): dZ = S*(G - <G,S>)/tau vanishes where the upstream G does,
// i.e. where (image, node) has no align coefficient and no pooled-gradient entry at an argmax row.  The kernels that
// build K5's inputs mark the blocks that CAN be nonzero; the dX / dW GEMMs skip the rest (gemm2_tc.cuh, `kact`):
//   t1[(row >> 8) * ld1 + (ccol >> 6)]   dX: A = dZ row tile of 256 x k-block of 64 compact columns
//   t2[(ccol >> 8) * ld2 + (row >> 6)]   dW: A = dZ^T column tile of 256 x k-block of 64 rows
// Data-dependent, conservative (a marked block may still be zero) and cleared by the forward's prologue launch.
struct DzBlockTables {
  uint8_t* t1; int ld1;
  uint8_t* t2; int ld2;
  const int32_t* pcol;        // [P] compact dZ column of a flat prototype (-1: none)
  // K5's own work items: iact[tile * iact_pitch + chunk], chunk = 32 locations of an image PAIR (image b of view 1 and
  // image b + imgs_first of view 2 share chunk b * cpi + location / 32); tile_of_node[n] = prototype tile of node n's
  // segment (-1: spill node, finished by row code).  iact == nullptr: not used
  uint8_t* iact; int iact_pitch;
  const int32_t* tile_of_node;
  int cpi;
};
// one row (view image v, location loc) x the whole column range [c_lo, c_hi] of node n: the softmax Jacobian couples all
// prototypes of a node, so a pooled-gradient entry of ONE prototype makes dZ nonzero in every column of its node at that row
__device__ __forceinline__ void mark_dz_entry(const DzBlockTables& b, int v, int loc, int HW, int imgs_first, int n, int c_lo,
                                              int c_hi) {
  const long long row = (long long)v * HW + loc;
  for (int c6 = c_lo >> 6; c6 <= (c_hi >> 6); ++c6) b.t1[(size_t)(row >> 8) * b.ld1 + c6] = 1;
  for (int c8 = c_lo >> 8; c8 <= (c_hi >> 8); ++c8) b.t2[(size_t)c8 * b.ld2 + (size_t)(row >> 6)] = 1;
  if (b.iact != nullptr) {
    const int tl = b.tile_of_node[n];
    if (tl >= 0) b.iact[(size_t)tl * b.iact_pitch + (size_t)(v >= imgs_first ? v - imgs_first : v) * b.cpi + (loc >> 5)] = 1;
  }
}
// the whole image pair (both views) x the node's columns: dense align gradient
__device__ __forceinline__ void mark_dz_image_node(const DzBlockTables& b, int img, int imgs_first, int V, int HW, int n, int c_lo,
                                                   int c_hi) {
  for (int view = 0; view < 2; ++view) {
    const int v = img + view * imgs_first;
    if (v >= V) break;
    const long long r_lo = (long long)v * HW, r_hi = r_lo + HW - 1;
    for (int c6 = c_lo >> 6; c6 <= (c_hi >> 6); ++c6)
      for (long long r8 = r_lo >> 8; r8 <= (r_hi >> 8); ++r8) b.t1[(size_t)r8 * b.ld1 + c6] = 1;
    for (int c8 = c_lo >> 8; c8 <= (c_hi >> 8); ++c8)
      for (long long r6 = r_lo >> 6; r6 <= (r_hi >> 6); ++r6) b.t2[(size_t)c8 * b.ld2 + (size_t)r6] = 1;
  }
  if (b.iact != nullptr) {
    const int tl = b.tile_of_node[n];
    if (tl >= 0)
      for (int k = 0; k < b.cpi; ++k) b.iact[(size_t)tl * b.iact_pitch + (size_t)img * b.cpi + k] = 1;
  }
}

// ---- head prologue: everything K1 needs, one launch ---------------------------------------------------------------
//   role A  pack the prototype kernels into their bf16 GEMM layouts (pack_weights_kernel's job)
//   role B  clear the packed max table and the align accumulators K1 merges into with atomics
//   role C  label tables (label_tables_kernel's job, without the n_desc memset: one warp per node counts)
struct PrologueParams {
  const float* w; const int32_t* row_map; int rows, C; __nv_bfloat16* wp;        // A (rows == 0: skipped)
  unsigned long long* packed; long long n_packed; double* align_sum; int n_align;   // B
  uint4* zero16; long long n_zero16;                                               // B: one more buffer to clear (16-byte units)
  const long long* ys; const int8_t* anc; int V, V_first, N, L;                     // C (ys == nullptr: skipped)
  int8_t* tgt; uint8_t* desc; int32_t* n_desc;
  int nb_pack, nb_zero, nb_tgt, nb_cnt;
};
__global__ void __launch_bounds__(256) head_prologue_kernel(const PrologueParams q) {
  int b = blockIdx.x;
  if (b < q.nb_pack) {
    const int c8 = q.C >> 3;
    const long long idx = (long long)b * 256 + threadIdx.x;
    if (idx >= (long long)q.rows * c8) return;
    const int r = int(idx / c8), c = int(idx - (long long)r * c8) * 8;
    const int src = q.row_map[r];
    uint4 o = make_uint4(0, 0, 0, 0);
    if (src >= 0) {
      const float4 a = *reinterpret_cast<const float4*>(q.w + (size_t)src * q.C + c);
      const float4 e = *reinterpret_cast<const float4*>(q.w + (size_t)src * q.C + c + 4);
      o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
      o.z = pack_bf16x2(e.x, e.y); o.w = pack_bf16x2(e.z, e.w);
    }
    *reinterpret_cast<uint4*>(q.wp + (size_t)r * q.C + c) = o;
    return;
  }
  b -= q.nb_pack;
  if (b < q.nb_zero) {                                  // 2 table entries (16 bytes) per thread; n_packed is even or the tail is scalar
    const long long i = ((long long)b * 256 + threadIdx.x) * 2;
    if (i + 1 < q.n_packed) *reinterpret_cast<uint4*>(q.packed + i) = make_uint4(0, 0, 0, 0);
    else if (i < q.n_packed) q.packed[i] = 0ull;
    if (b == 0 && q.align_sum != nullptr)
      for (int n = threadIdx.x; n < q.n_align; n += 256) q.align_sum[n] = 0.0;
    for (long long z = (long long)b * 256 + threadIdx.x; z < q.n_zero16; z += (long long)q.nb_zero * 256)
      q.zero16[z] = make_uint4(0, 0, 0, 0);
    return;
  }
  b -= q.nb_zero;
  if (b < q.nb_tgt) {
    const int idx = b * 256 + threadIdx.x;
    if (idx >= q.V * q.N) return;
    const int v = idx / q.N, n = idx - v * q.N;
    const long long y = q.ys[v];
    const int8_t t = (y >= 0 && y < q.L) ? q.anc[(size_t)y * q.N + n] : int8_t(-1);
    q.tgt[idx] = t;
    if (v < q.V_first) q.desc[idx] = (t >= 0);
    return;
  }
  b -= q.nb_tgt;
  {                                                     // n_desc[n]: one warp per node, lanes over the rows
    const int n = b * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= q.N) return;
    int cnt = 0;
    for (int v = lane; v < q.V; v += 32) {
      const long long y = q.ys[v];
      cnt += (y >= 0 && y < q.L && q.anc[(size_t)y * q.N + n] >= 0) ? 1 : 0;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) q.n_desc[n] = cnt;
  }
}

// ---- forward finish: pool unpack + align finalize + classifier, one launch ----------------------------------------
// (unpack_pool_kernel + align_finalize_kernel + classifier_fwd_kernel; the classifier reads the packed table itself, so
// the three roles are independent)
struct PoolClassifyParams {
  const unsigned long long* packed; const double* align_sum; const int32_t* n_desc;
  const float* wc; const float* bias;
  const int32_t *col_node, *proto_off, *cls_off, *wc_off;
  int V, P, K, N, HW; float thresh;
  float* pooled; int32_t* argmax; float* align; float* out;
  int nb_unpack, nb_cls, nb_align;
  // deferred riders (narrow spill nodes, see spill_nodes.cuh): the fused forward only wrote their raw logits; one block
  // per (rider, image pair) finishes softmax / max-pool / align for the pair's locations, then unpacks the pair's pooled
  // entries of that node and applies its classifier; the last block of a rider finalizes the node's align loss
  int n_riders;
  SpillParams rider[4];
  double* align_sum_rw;
  unsigned int* counter;     // [1 + rider]: zero on entry, left zero
};
// Rider role of the forward finish kernel: ONE block (512 threads) per (rider, image pair).  Same arithmetic as the
// fused kernel's epilogue (softmax_row: 4-way interleaved sums, ex2.approx, rcp.approx; align inner product in the same
// association order), organised for a plain SIMT kernel with few registers:
//   phase 1  thread per location: row maximum and 1 / row sum of both views -> shared memory; align term of the pair row
//   phase 2  thread per (view, prototype, row slice): walks its slice in location order, first maximum wins
//   phase 3  thread per (view, prototype): best of the slices in order -> pooled / argmax (written directly, the packed
//            table is not used for rider columns); then the node's classifier rows for the two images
// The last block of a rider finalizes the node's align loss.
__device__ __forceinline__ void rider_finish_block(const PoolClassifyParams& q, const SpillParams& sp, int r, int b, float* rsm) {
  const int HW = q.HW, Pn = sp.P_n, tid = threadIdx.x, nt = blockDim.x;
  const int st = Pn | 1;                                    // odd row pitch: conflict-free column and row walks
  float* zt = rsm;                                          // [2][HW][st] the pair's raw logits, staged once
  float* mk1 = zt + (size_t)2 * HW * st;
  float* inv1 = mk1 + HW;
  float* mk2 = inv1 + HW;
  float* inv2 = mk2 + HW;
  float* bval = inv2 + HW;                                  // [512]
  int* bloc = reinterpret_cast<int*>(bval + 512);           // [512]
  float* pooled_s = reinterpret_cast<float*>(bloc + 512);   // [2][64]
  float* red = pooled_s + 128;                              // [32]
  const int v1 = b, v2 = b + sp.imgs_first;
  const bool has2 = v2 < q.V;
  const float sc = sp.scale_log2;
  const size_t rowA = (size_t)b * HW, rowB = (size_t)sp.halfM + (size_t)b * HW;
  const bool use_align = sp.desc != nullptr && sp.align_sum != nullptr && has2 && sp.desc[(size_t)b * sp.n_nodes + sp.node] != 0;
  float a_sum = 0.f;
  {                                                         // stage: 16-byte coalesced loads, all independent
    const int n4 = (Pn + 3) >> 2;
    const int per_view = HW * n4, total = (has2 ? 2 : 1) * per_view;
    for (int base = tid; base < total; base += 8 * nt) {        // eight loads in flight per thread before the first store
      float4 t[8];
      int dst[8], qd8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = base + u * nt;
        dst[u] = -1;
        if (idx < total) {
          const int view = idx >= per_view, rem = idx - view * per_view;
          const int row = rem / n4, qd = rem - row * n4;
          t[u] = __ldg(reinterpret_cast<const float4*>(sp.zs + ((view ? rowB : rowA) + row) * sp.ldz + sp.zoff) + qd);
          dst[u] = (view * HW + row) * st + 4 * qd;
          qd8[u] = qd;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (dst[u] >= 0) {
          float* d = zt + dst[u];
          d[0] = t[u].x;
          if (4 * qd8[u] + 1 < Pn) d[1] = t[u].y;
          if (4 * qd8[u] + 2 < Pn) d[2] = t[u].z;
          if (4 * qd8[u] + 3 < Pn) d[3] = t[u].w;
        }
      }
    }
  }
  __syncthreads();
  for (int row = tid; row < HW; row += nt) {
    const float* z1 = zt + (size_t)row * st;
    const float* z2 = zt + ((size_t)HW + row) * st;
    float m1 = -INFINITY, m2 = -INFINITY;
    for (int i = 0; i < Pn; ++i) {
      m1 = fmaxf(m1, z1[i]);
      if (has2) m2 = fmaxf(m2, z2[i]);
    }
    const float k1 = __fmul_rn(m1, sc), k2 = __fmul_rn(m2, sc);
    float l1[4] = {0.f, 0.f, 0.f, 0.f}, l2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i0 = 0; i0 < Pn; i0 += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u;
        if (i < Pn) {
          l1[u] = __fadd_rn(l1[u], ex2(__fmaf_rn(z1[i], sc, -k1)));
          if (has2) l2[u] = __fadd_rn(l2[u], ex2(__fmaf_rn(z2[i], sc, -k2)));
        }
      }
    }
    const float i1 = rcp_approx(__fadd_rn(__fadd_rn(l1[0], l1[1]), __fadd_rn(l1[2], l1[3])));
    const float i2 = has2 ? rcp_approx(__fadd_rn(__fadd_rn(l2[0], l2[1]), __fadd_rn(l2[2], l2[3]))) : 0.f;
    mk1[row] = k1; inv1[row] = i1; mk2[row] = k2; inv2[row] = i2;
    if (use_align) {
      float ip4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i0 = 0; i0 < Pn; i0 += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u;
          if (i < Pn) {
            const float s1 = __fmul_rn(ex2(__fmaf_rn(z1[i], sc, -k1)), i1);
            const float s2 = __fmul_rn(ex2(__fmaf_rn(z2[i], sc, -k2)), i2);
            ip4[u] = __fmaf_rn(s1, s2, ip4[u]);
          }
        }
      }
      const float ip = __fadd_rn(__fadd_rn(ip4[0], ip4[1]), __fadd_rn(ip4[2], ip4[3]));
      a_sum += -__logf(ip + 1e-12f);
    }
  }
  __syncthreads();
  int nsl = nt / (2 * Pn);
  if (nsl > 16) nsl = 16;
  if (nsl < 1) nsl = 1;
  const int per = (HW + nsl - 1) / nsl;
  if (tid < 2 * Pn * nsl) {
    const int sl = tid / (2 * Pn), vc = tid - sl * 2 * Pn;
    const int view = vc >= Pn, c = vc - view * Pn;
    float best = -1.f;
    int arg = 0;
    if (view == 0 || has2) {
      const float* z = zt + (size_t)view * HW * st + c;
      const float* mk = view ? mk2 : mk1;
      const float* iv = view ? inv2 : inv1;
      const int r0 = sl * per, r1 = min(HW, r0 + per);
#pragma unroll 4
      for (int row = r0; row < r1; ++row) {
        const float sv = __fmul_rn(ex2(__fmaf_rn(z[(size_t)row * st], sc, -mk[row])), iv[row]);
        if (sv > best) { best = sv; arg = row; }
      }
    }
    bval[tid] = best;
    bloc[tid] = arg;
  }
  __syncthreads();
  if (tid < 2 * Pn) {
    const int view = tid >= Pn, c = tid - view * Pn;
    if (view == 0 || has2) {
      float best = -1.f;
      int arg = 0;
      for (int sl = 0; sl < nsl; ++sl) {
        const float x = bval[sl * 2 * Pn + tid];
        if (x > best) { best = x; arg = bloc[sl * 2 * Pn + tid]; }
      }
      if (best < q.thresh) best = 0.f;
      const size_t i = (size_t)(view ? v2 : v1) * q.P + sp.poff + c;
      q.pooled[i] = best;
      q.argmax[i] = arg;
      pooled_s[view * 64 + c] = best;
    }
  }
  __syncthreads();
  if (q.out != nullptr) {
    const int k0 = q.cls_off[sp.node], kn = q.cls_off[sp.node + 1] - k0;
    for (int idx = tid; idx < 2 * kn; idx += nt) {
      const int view = idx >= kn, c = idx - view * kn;
      if (view == 1 && !has2) continue;
      const float* w = q.wc + q.wc_off[sp.node] + (size_t)c * Pn;
      float acc = 0.f;
      for (int pp = 0; pp < Pn; ++pp) acc = fmaf(fmaxf(w[pp], 0.f), pooled_s[view * 64 + pp], acc);
      q.out[(size_t)(view ? v2 : v1) * q.K + k0 + c] = acc + (q.bias ? q.bias[k0 + c] : 0.f);
    }
  }
  if (q.align != nullptr) {                             // block sum of the align terms, then last block: the node's loss
    const float tot = block_sum(a_sum, red);
    if (tid == 0) {
      if (tot != 0.f) atomicAdd(q.align_sum_rw + sp.node, (double)tot);
      __threadfence();
      const bool last = atomicAdd(q.counter + 1 + r, 1u) == (unsigned)sp.imgs_first - 1u;
      if (last) {
        __threadfence();
        const int nd = q.n_desc[sp.node] / 2;
        q.align[sp.node] = nd > 0 ? float(__ldcg(q.align_sum_rw + sp.node) / (double(nd) * double(HW))) : 0.f;
        q.counter[1 + r] = 0u;
      }
    }
  }
}

// RIDERS = true carries the rider role (its row code needs ~220 registers: one block per SM, so the unpack role of that
// variant strides over the table with few blocks); layouts without deferred riders run the lean variant.
template <bool RIDERS>
__global__ void __launch_bounds__(RIDERS ? 512 : 256) pool_classify_fwd_kernel(const PoolClassifyParams q) {
  constexpr int BS = RIDERS ? 512 : 256;
  int b = blockIdx.x;
  if (b < q.nb_unpack) {
    const long long n = (long long)q.V * q.P;
    for (long long i = (long long)b * BS + threadIdx.x; i < n; i += (long long)q.nb_unpack * BS) {
      if constexpr (RIDERS) {                           // rider columns are written by their rider blocks
        const int pcol = int(i % q.P);
        bool skip = false;
        for (int r = 0; r < q.n_riders; ++r) skip |= pcol >= q.rider[r].poff && pcol < q.rider[r].poff + q.rider[r].P_n;
        if (skip) continue;
      }
      const unsigned long long k = q.packed[i];
      float v = __uint_as_float((uint32_t)(k >> 32));
      if (v < q.thresh) v = 0.f;
      q.pooled[i] = v;
      q.argmax[i] = (int32_t)(0xFFFFFFFFu - (uint32_t)k);
    }
    return;
  }
  b -= q.nb_unpack;
  if (b < q.nb_cls) {
    const int idx = b * BS + threadIdx.x;
    if (idx >= q.V * q.K) return;
    const int v = idx / q.K, k = idx - v * q.K;
    const int n = q.col_node[k];
    if constexpr (RIDERS)
      for (int r = 0; r < q.n_riders; ++r)
        if (n == q.rider[r].node) return;
    const int p0 = q.proto_off[n], pn = q.proto_off[n + 1] - p0;
    const float* w = q.wc + q.wc_off[n] + (size_t)(k - q.cls_off[n]) * pn;
    const unsigned long long* x = q.packed + (size_t)v * q.P + p0;
    float acc = 0.f;
    for (int p = 0; p < pn; ++p) {
      float xv = __uint_as_float((uint32_t)(x[p] >> 32));
      if (xv < q.thresh) xv = 0.f;
      acc = fmaf(fmaxf(w[p], 0.f), xv, acc);
    }
    q.out[idx] = acc + (q.bias ? q.bias[k] : 0.f);
    return;
  }
  b -= q.nb_cls;
  if (b < q.nb_align) {
    const int n = b * BS + threadIdx.x;
    if (n >= q.N) return;
    if constexpr (RIDERS)
      for (int r = 0; r < q.n_riders; ++r)
        if (n == q.rider[r].node) return;
    const int nd = q.n_desc[n] / 2;
    q.align[n] = nd > 0 ? float(q.align_sum[n] / (double(nd) * double(q.HW))) : 0.f;
    return;
  }
  if constexpr (RIDERS) {
    extern __shared__ float rsm[];
    b -= q.nb_align;
    const int r = b / q.rider[0].imgs_first, img = b - r * q.rider[0].imgs_first;
    if (r < q.n_riders) rider_finish_block(q, q.rider[r], r, img, rsm);
  }
}

// ---- all per-node loss terms + their combination, one launch ------------------------------------------------------
// grid (N, 3 or 4): y = 0 / 1 the tanh term of a view half, y = 2 the class term (+ accuracy counters, + the row
// log-sum-exp kept for the backward), y = 3 ||E_n||^2 of the orth term (orth_sumsq_kernel's job; E from hcomp_orth_gram);
// the LAST block to finish (device counter, left at zero) combines the terms.
struct ChainFwdParams {
  const float* pooled; const float* out; const float* align; float* orth_sq;   // align / orth_sq may be NULL
  const float* E; int P_max;                                                    // Gram slab [N, P_max, P_max] (orth on)
  const int8_t* tgt; const int32_t* n_desc; const float* child_w;
  const int32_t *proto_off, *cls_off;
  int V, V_first, N, P, K; float eps, mult; int do_tanh, do_cls;
  LossWeights lw;
  float *tanh_part, *colsum, *cls, *lse; int32_t* n_correct;
  float* stats; float* total; unsigned int* counter;
};
__global__ void __launch_bounds__(256) head_chain_fwd_kernel(const ChainFwdParams q) {
  __shared__ float sh[8][64];
  __shared__ float red[32];
  __shared__ int last;
  const int n = blockIdx.x, role = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (role < 2) {
    if (q.do_tanh) {                                   // same arithmetic as tanh_loss_fwd_kernel
      const int h = role;
      const int p0 = q.proto_off[n], pn = q.proto_off[n + 1] - p0;
      const int vb = h ? q.V_first : 0, ve = h ? q.V : q.V_first;
      float acc = 0.f;
      for (int pc = 0; pc < pn; pc += 64) {
        float t0 = 0.f, t1 = 0.f;
        const int pa = pc + lane, pb = pc + 32 + lane;
        // four rows per trip, every load issued before the first use (the kernel is pure load latency); same summation
        // order as tanh_loss_fwd_kernel (rows ascending per warp)
        for (int v = vb + warp; v < ve; v += 32) {
          int on[4];
          float x0[4], x1[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int vv = v + 8 * u;
            const bool ok = vv < ve;
            on[u] = ok ? int(q.tgt[(size_t)vv * q.N + n]) : -1;
            const float* row = q.pooled + (size_t)(ok ? vv : v) * q.P + p0;
            x0[u] = (pa < pn) ? row[pa] : 0.f;
            x1[u] = (pb < pn) ? row[pb] : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (on[u] >= 0) { t0 += x0[u]; t1 += x1[u]; }
        }
        sh[warp][lane] = t0;
        sh[warp][32 + lane] = t1;
        __syncthreads();
        if (threadIdx.x < 64 && pc + threadIdx.x < pn) {
          float t = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
          q.colsum[(size_t)h * q.P + p0 + pc + threadIdx.x] = t;
          acc += logf(tanhf(t) + q.eps);
        }
        __syncthreads();
      }
      acc = block_sum(acc, red);
      if (threadIdx.x == 0) q.tanh_part[h * q.N + n] = q.n_desc[n] > 0 ? -0.5f * acc / float(pn) : 0.f;
    }
  } else if (role == 3) {                              // ||E_n||_F^2, fixed order (deterministic)
    const int pn = q.proto_off[n + 1] - q.proto_off[n];
    const float* En = q.E + (size_t)n * q.P_max * q.P_max;
    float acc = 0.f;
    for (int ij = threadIdx.x; ij < pn * pn; ij += 256) {
      const float e = En[(ij / pn) * q.P_max + (ij % pn)];
      acc = fmaf(e, e, acc);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) q.orth_sq[n] = acc;
  } else {                                             // class term: class_loss_fwd_kernel + lse
    const int k0 = q.cls_off[n], kn = q.cls_off[n + 1] - k0;
    float acc = 0.f, corr = 0.f;
    for (int v = threadIdx.x; v < q.V; v += 256) {
      const int t = q.tgt[(size_t)v * q.N + n];
      if (t < 0) continue;
      const float* o = q.out + (size_t)v * q.K + k0;
      float mx = -INFINITY, best = -INFINITY;
      int arg = 0;
      for (int c = 0; c < kn; ++c) {
        const float x = sparsity_x(o[c], q.mult);
        mx = fmaxf(mx, x);
        if (o[c] > best) { best = o[c]; arg = c; }
      }
      float se = 0.f, xt = 0.f;
      for (int c = 0; c < kn; ++c) {
        const float x = sparsity_x(o[c], q.mult);
        se += expf(x - mx);
        if (c == t) xt = x;
      }
      const float l = logf(se) + mx;
      q.lse[(size_t)v * q.N + n] = l;
      acc += q.child_w[k0 + t] * (l - xt);
      corr += (arg == t) ? 1.f : 0.f;
    }
    acc = block_sum(acc, red);
    corr = block_sum(corr, red);
    if (threadIdx.x == 0) {
      const int nd = q.n_desc[n];
      q.cls[n] = nd > 0 ? acc / float(nd) : 0.f;
      q.n_correct[n] = int(corr + 0.5f);
    }
  }
  // ---- last block done: combine (loss_combine_kernel's arithmetic)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(q.counter, 1u) == gridDim.x * gridDim.y - 1u);
  __syncthreads();
  if (!last) return;
  __threadfence();
  float acc = 0.f;
  for (int m = threadIdx.x; m < q.N; m += 256) {
    const bool on = q.n_desc[m] > 0;
    const float a = (q.align && on) ? __ldcg(q.align + m) : 0.f;
    const float t = (q.do_tanh && on) ? __ldcg(q.tanh_part + m) + __ldcg(q.tanh_part + q.N + m) : 0.f;
    const float o = (q.orth_sq && on) ? sqrtf(__ldcg(q.orth_sq + m)) : 0.f;
    const float c = (q.do_cls && on) ? __ldcg(q.cls + m) : 0.f;
    q.stats[m] = a; q.stats[q.N + m] = t; q.stats[2 * q.N + m] = o; q.stats[3 * q.N + m] = c;
    acc += q.lw.w[0] * a + q.lw.w[1] * t + q.lw.w[2] * o + q.lw.w[3] * c;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) { *q.total = acc; *q.counter = 0u; }
}

// ---- backward of the tanh + class terms chained through the classifier, one launch --------------------------------
//   role A  g_pooled[v,p] = tanh term + sum_c g_out[v,c] * relu(Wc[c,p])          (tanh_loss_bwd + class_loss_bwd +
//   role B  g_Wc[c,p]     = [Wc > 0] * sum_v g_out[v,c] * pooled[v,p]              classifier_bwd_* + the autograd add)
//   role C  g_bias[k]     = sum_v g_out[v,k]
//   role D  g_align[n]    = g_total * w_align
// with g_out[v,c] = g_total * w_class / n_desc * w[t] * (softmax(x)[c] - [c == t]) * dx/dout rebuilt on the fly from the
// logits and the row log-sum-exp of the forward (never stored).
struct ChainBwdParams {
  const float* g_total; const float* pooled; const float* out; const float* wc; const float* colsum; const float* lse;
  const int8_t* tgt; const int32_t* n_desc; const float* child_w;
  const int32_t *proto_node, *proto_off, *cls_off, *wc_off, *col_node, *welem_col, *welem_proto;
  int V, V_first, N, P, K, n_w; float eps, mult; int do_tanh, do_cls;
  LossWeights lw;
  float *g_pooled, *g_wc, *g_bias, *g_align;            // any may be NULL
  int nb_pooled, nb_wc, nb_bias, nb_align;
  // optional: K5's scatter table / align coefficients straight from here (bwd_prep_kernel's job), valid when this
  // g_pooled / g_align reach the head backward unchanged (the host checks)
  const int32_t* argmax; float thresh; int2* scat; const uint8_t* desc; int HW; float* coef;
  DzBlockTables blk;          // blk.t1 == nullptr: no block tables
};
__device__ __forceinline__ float chain_gout(const ChainBwdParams& q, int v, int n, int t, int k0, int c, float coef) {
  const float o = q.out[(size_t)v * q.K + k0 + c];
  const float sm = expf(sparsity_x(o, q.mult) - q.lse[(size_t)v * q.N + n]);
  return coef * (sm - (c == t ? 1.f : 0.f)) * sparsity_dx(o, q.mult);
}
__global__ void __launch_bounds__(256) head_chain_bwd_kernel(const ChainBwdParams q) {
  int b = blockIdx.x;
  const float gT = q.g_total[0];
  if (b < q.nb_pooled) {
    const long long idx = (long long)b * 256 + threadIdx.x;
    if (idx >= (long long)q.V * q.P) return;
    const int v = int(idx / q.P), p = int(idx - (long long)v * q.P);
    const int n = q.proto_node[p];
    const int t = q.tgt[(size_t)v * q.N + n];
    float g = 0.f;
    if (t >= 0) {
      const int p0 = q.proto_off[n], pn = q.proto_off[n + 1] - p0;
      if (q.do_tanh) {
        const float th = tanhf(q.colsum[(size_t)(v >= q.V_first) * q.P + p]);
        g = gT * q.lw.w[1] * (-0.5f / float(pn)) * (1.f - th * th) / (th + q.eps);
      }
      if (q.do_cls) {
        const int k0 = q.cls_off[n], kn = q.cls_off[n + 1] - k0;
        const float coef = gT * q.lw.w[3] / float(q.n_desc[n]) * q.child_w[k0 + t];
        const float* w = q.wc + q.wc_off[n] + (p - p0);
        float acc = 0.f;
#pragma unroll 4
        for (int c = 0; c < kn; ++c) acc = fmaf(chain_gout(q, v, n, t, k0, c, coef), fmaxf(w[(size_t)c * pn], 0.f), acc);
        g += acc;
      }
    }
    q.g_pooled[idx] = g;
    if (q.scat != nullptr) {
      const float gs = (q.thresh > 0.f && q.pooled[idx] < q.thresh) ? 0.f : g;    // inference threshold kills the gradient too
      const int am = q.argmax[idx];
      q.scat[idx] = make_int2(am, __float_as_int(gs));
      if (q.blk.t1 != nullptr && gs != 0.f) {
        const int ca = q.blk.pcol[q.proto_off[n]], cb = q.blk.pcol[q.proto_off[n + 1] - 1];
        if (ca >= 0 && cb >= ca) mark_dz_entry(q.blk, v, am, q.HW, q.V_first, n, ca, cb);
      }
    }
    return;
  }
  b -= q.nb_pooled;
  if (b < q.nb_wc) {                                    // one warp per classifier weight element
    const int wid = b * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (wid >= q.n_w) return;
    float acc = 0.f;
    if (q.do_cls && q.wc[wid] > 0.f) {
      const int k = q.welem_col[wid], p = q.welem_proto[wid];
      const int n = q.col_node[k];
      const int k0 = q.cls_off[n];
      const int nd = q.n_desc[n];
      const float base = nd > 0 ? gT * q.lw.w[3] / float(nd) : 0.f;
#pragma unroll 4
      for (int v = lane; v < q.V; v += 32) {            // loads unconditional (rows without a label read a valid dummy
        const int t = q.tgt[(size_t)v * q.N + n];       // column; their lse is uninitialised, the select drops the result)
        const float gout = chain_gout(q, v, n, t, k0, k - k0, base * q.child_w[k0 + (t < 0 ? 0 : t)]);
        const float pl = q.pooled[(size_t)v * q.P + p];
        acc = t >= 0 ? fmaf(gout, pl, acc) : acc;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) q.g_wc[wid] = acc;
    return;
  }
  b -= q.nb_wc;
  if (b < q.nb_bias) {
    const int k = b * 256 + threadIdx.x;
    if (k >= q.K) return;
    float acc = 0.f;
    if (q.do_cls) {
      const int n = q.col_node[k];
      const int k0 = q.cls_off[n];
      const int nd = q.n_desc[n];
      const float base = nd > 0 ? gT * q.lw.w[3] / float(nd) : 0.f;
      for (int v = 0; v < q.V; ++v) {
        const int t = q.tgt[(size_t)v * q.N + n];
        if (t >= 0) acc += chain_gout(q, v, n, t, k0, k - k0, base * q.child_w[k0 + t]);
      }
    }
    q.g_bias[k] = acc;
    return;
  }
  b -= q.nb_bias;
  if (b < q.nb_align) {
    const int n = b * 256 + threadIdx.x;
    if (n < q.N) q.g_align[n] = gT * q.lw.w[0];
    return;
  }
  b -= q.nb_align;
  const int idx = b * 256 + threadIdx.x;                // align coefficients [V_first, N] (align_coef_kernel)
  if (idx >= q.V_first * q.N) return;
  const int m = idx % q.N;
  const int nd = q.n_desc[m] / 2;
  const float cf = (q.desc[idx] && nd > 0) ? gT * q.lw.w[0] * 0.5f / (float(nd) * float(q.HW)) : 0.f;
  q.coef[idx] = cf;
  if (q.blk.t1 != nullptr && cf != 0.f) {
    const int pa = q.proto_off[m], pb = q.proto_off[m + 1] - 1;
    const int ca = q.blk.pcol[pa], cb = q.blk.pcol[pb];
    if (ca >= 0 && cb >= ca) mark_dz_image_node(q.blk, idx / q.N, q.V_first, q.V, q.HW, m, ca, cb);
  }
}

// orth_bwd_kernel with the upstream gradient taken as g_total * weight (no loss_grads launch in front of it)
__global__ void orth_bwd_scaled_kernel(const float* __restrict__ w, const int32_t* __restrict__ proto_node,
                                       const int32_t* __restrict__ proto_off, int C, int P_max, const float* __restrict__ loss,
                                       const float* __restrict__ E, const uint8_t* __restrict__ rel,
                                       const float* __restrict__ g_total, float weight, float* __restrict__ g_w) {
  const int row = blockIdx.x;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int n = proto_node[row];
  const int p0 = proto_off[n], pn = proto_off[n + 1] - p0;
  const float L = loss[n], g = g_total[0] * weight;
  float acc = 0.f;
  if (rel[row] && L > 0.f && g != 0.f) {
    const float* Er = E + ((size_t)n * P_max + (row - p0)) * P_max;
    const float* wn = w + (size_t)p0 * C + c;
#pragma unroll 4
    for (int j = 0; j < pn; ++j) acc = fmaf(Er[j], wn[(size_t)j * C], acc);
    acc *= g * 2.f / L;
  }
  g_w[(size_t)row * C + c] = acc;
}

// ---- backward prep: scatter table + align coefficients, one launch (make_scat_kernel + align_coef_kernel) ---------
__global__ void __launch_bounds__(256) bwd_prep_kernel(const int32_t* __restrict__ argmax, const float* __restrict__ g_pooled,
                                                       const float* __restrict__ pooled, float thresh, long long n,
                                                       int2* __restrict__ scat, int nb_scat, const uint8_t* __restrict__ desc,
                                                       const int32_t* __restrict__ n_desc, const float* __restrict__ g_align,
                                                       int B, int N, int HW, float* __restrict__ coef, int P, int V,
                                                       const int32_t* __restrict__ proto_off,
                                                       const int32_t* __restrict__ proto_node, const DzBlockTables blk) {
  int b = blockIdx.x;
  if (b < nb_scat) {
    const long long i = (long long)b * 256 + threadIdx.x;
    if (i >= n) return;
    float g = g_pooled[i];
    if (pooled != nullptr && pooled[i] < thresh) g = 0.f;
    const int am = argmax[i];
    scat[i] = make_int2(am, __float_as_int(g));
    if (blk.t1 != nullptr && g != 0.f) {
      const int v = int(i / P), pp = int(i - (long long)v * P);
      const int nd = proto_node[pp];
      const int ca = blk.pcol[proto_off[nd]], cb = blk.pcol[proto_off[nd + 1] - 1];
      if (ca >= 0 && cb >= ca) mark_dz_entry(blk, v, am, HW, B, nd, ca, cb);
    }
    return;
  }
  b -= nb_scat;
  const int idx = b * 256 + threadIdx.x;
  if (idx >= B * N) return;
  const int m = idx % N;
  const int nd = n_desc[m] / 2;
  const float cf = (desc[idx] && nd > 0) ? g_align[m] * 0.5f / (float(nd) * float(HW)) : 0.f;
  coef[idx] = cf;
  if (blk.t1 != nullptr && cf != 0.f) {
    const int pa = proto_off[m], pb = proto_off[m + 1] - 1;
    const int ca = blk.pcol[pa], cb = blk.pcol[pb];
    if (ca >= 0 && cb >= ca) mark_dz_image_node(blk, idx / N, B, V, HW, m, ca, cb);
  }
}

// ---------------------------------------------------------------- joint leaf distribution (util/node.py:383-385, pipnet/pipnet.py:173-185)
// probs[v,k] = softmax_c(log1p(out^2)/tau) within each node; leaf[v,l] = product of probs along the path.
// override[k] >= 0 forces the probability of child column k for every sample (leave-out classes: 1 on the left-out leaf
// child, 0 on its siblings, util/node.py:319-323; fully masked class under the overspecificity mask: leaf-count
// fractions, :335-359); a node is overridden as a whole.
__global__ void node_probs_kernel(const float* __restrict__ out, const int32_t* __restrict__ cls_off, int V, int N, int K,
                                  float inv_tau, const float* __restrict__ override, float* __restrict__ probs) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= V * N) return;
  const int v = idx / N, n = idx - v * N;
  const int k0 = cls_off[n], kn = cls_off[n + 1] - k0;
  if (override != nullptr && override[k0] >= 0.f) {
    for (int c = 0; c < kn; ++c) probs[(size_t)v * K + k0 + c] = override[k0 + c];
    return;
  }
  const float* o = out + (size_t)v * K + k0;
  float mx = -INFINITY;
  for (int c = 0; c < kn; ++c) mx = fmaxf(mx, log1pf(o[c] * o[c]) * inv_tau);
  float se = 0.f;
  for (int c = 0; c < kn; ++c) se += expf(log1pf(o[c] * o[c]) * inv_tau - mx);
  for (int c = 0; c < kn; ++c) probs[(size_t)v * K + k0 + c] = expf(log1pf(o[c] * o[c]) * inv_tau - mx) / se;
}
__global__ void leaf_joint_kernel(const float* __restrict__ probs, const int32_t* __restrict__ path_off,
                                  const int32_t* __restrict__ path_col, int V, int L, int K, float* __restrict__ joint) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= V * L) return;
  const int v = idx / L, l = idx - v * L;
  float pr = 1.f;
  for (int i = path_off[l]; i < path_off[l + 1]; ++i) pr *= probs[(size_t)v * K + path_col[i]];
  joint[idx] = pr;
}
__global__ void row_argmax_kernel(const float* __restrict__ x, int V, int L, long long* __restrict__ pred) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  float best = -INFINITY;
  int arg = 0;
  for (int l = 0; l < L; ++l) {
    const float t = x[(size_t)v * L + l];
    if (t > best) { best = t; arg = l; }
  }
  pred[v] = arg;
}

// ---------------------------------------------------------------- full softmax map of ONE node (visualisation path)
// Rebuilds S_n[v, p, hw] for a single node in fp32 from bf16 features (util/vis_hpipnet.py:62-127 reads it
// at batch size 1); one warp per (v, hw).
__global__ void materialize_map_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, int V, int HW,
                                       int C, int pn, float inv_tau, float* __restrict__ map /* [V, pn, HW] */) {
  extern __shared__ float zbuf[];   // [warps][pn]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const long long r = (long long)blockIdx.x * nw + warp;
  if (r >= (long long)V * HW) return;
  const int v = int(r / HW), hw = int(r - (long long)v * HW);
  float* z = zbuf + warp * pn;
  const __nv_bfloat16* xr = x + (size_t)r * C;
  for (int p = 0; p < pn; ++p) {
    float d = 0.f;
    for (int c = lane; c < C; c += 32) d = fmaf(__bfloat162float(xr[c]), w[(size_t)p * C + c], d);
    d = warp_sum(d);
    if (lane == 0) z[p] = d * inv_tau;
  }
  __syncwarp();
  float mx = -INFINITY;
  for (int p = 0; p < pn; ++p) mx = fmaxf(mx, z[p]);
  float se = 0.f;
  for (int p = 0; p < pn; ++p) se += expf(z[p] - mx);
  for (int p = lane; p < pn; p += 32) map[((size_t)v * pn + p) * HW + hw] = expf(z[p] - mx) / se;
}

}  // namespace hc
