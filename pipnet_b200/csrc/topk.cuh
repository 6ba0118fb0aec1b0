// Visualisation feed: per-(prototype, leaf) top-k activations with image id and argmax location, updated
// from the pooled scores / argmax the fused forward already streams out -- for ALL nodes in one pass over the data with
// any batch size.  The reference (util/vis_hpipnet.py:184-290, save_images_topk) re-runs the whole network per node with
// batch size 1 and keeps Python heaps per (prototype, leaf); which images may enter a heap follows its rules:
//   * only images whose leaf lies below the prototype's node (ModifiedLabelLoader filtering),
//   * prototypes without any relevant class (classifier column > 1e-3) are skipped (:243-244),
//   * the image's child class at that node must be relevant to the prototype (:250), or must NOT be when
//     find_non_descendants is set (:264).
// One thread per (prototype, leaf of the batch): it walks the batch rows of that leaf through the linked list built by
// desc_prep_kernel and inserts into the sorted k-list, so no two threads ever touch the same list.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hc {

constexpr int TOPK_MAX = 32;

__global__ void topk_update_kernel(const float* __restrict__ pooled, const int32_t* __restrict__ argmax,
                                   const long long* __restrict__ ys, const long long* __restrict__ img_ids,
                                   const int32_t* __restrict__ leader, const int32_t* __restrict__ next,
                                   const int8_t* __restrict__ anc, const int32_t* __restrict__ proto_node,
                                   const int32_t* __restrict__ proto_off, const int32_t* __restrict__ cls_off,
                                   const int32_t* __restrict__ wc_off, const float* __restrict__ wc, int find_non_desc,
                                   int V, int P, int N, int L, int k, float* __restrict__ t_score,
                                   long long* __restrict__ t_img, int32_t* __restrict__ t_loc) {
  const int r = blockIdx.y;
  if (leader[r] != r) return;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const long long leaf = ys[r];
  if (leaf < 0 || leaf >= L) return;
  const int n = proto_node[p];
  const int c = anc[(size_t)leaf * N + n];
  if (c < 0) return;                                        // leaf not below this prototype's node
  const int p0 = proto_off[n], pn = proto_off[n + 1] - p0, pl = p - p0;
  const int cn = cls_off[n + 1] - cls_off[n];
  bool any_rel = false, own_rel = false;
  for (int cc = 0; cc < cn; ++cc) {
    const bool rel = wc[wc_off[n] + (size_t)cc * pn + pl] > 1e-3f;
    any_rel |= rel;
    if (cc == c) own_rel = rel;
  }
  if (!any_rel || (find_non_desc ? own_rel : !own_rel)) return;
  float sc[TOPK_MAX];
  long long im[TOPK_MAX];
  int32_t lo[TOPK_MAX];
  const size_t base = ((size_t)p * L + (size_t)leaf) * k;
  for (int i = 0; i < k; ++i) { sc[i] = t_score[base + i]; im[i] = t_img[base + i]; lo[i] = t_loc[base + i]; }
  for (int q = r; q >= 0; q = next[q]) {
    const float s = pooled[(size_t)q * P + p];
    // list is sorted by descending score, empty slots (img = -1) at the end; equal scores keep the earlier image
    if (im[k - 1] >= 0 && !(s > sc[k - 1])) continue;
    int pos = k - 1;
    while (pos > 0 && (im[pos - 1] < 0 || s > sc[pos - 1])) {
      sc[pos] = sc[pos - 1]; im[pos] = im[pos - 1]; lo[pos] = lo[pos - 1];
      --pos;
    }
    sc[pos] = s; im[pos] = img_ids[q]; lo[pos] = argmax[(size_t)q * P + p];
  }
  for (int i = 0; i < k; ++i) { t_score[base + i] = sc[i]; t_img[base + i] = im[i]; t_loc[base + i] = lo[i]; }
}

}  // namespace hc
