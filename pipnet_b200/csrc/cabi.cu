// C-ABI entry points of libhcomp_head.so (declared in include/hcomp_head.h).
// Host side only: argument checks, TMA tensor-map encoding, launch geometry.  No allocation,
// no stream synchronisation, no CPU fallback.
#include "../../include/hcomp_head.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>

#include "gemm_tc.cuh"
#include "gemm2_tc.cuh"
#include "head_pair.cuh"
#include "small_kernels.cuh"
#include "desc_losses.cuh"
#include "topk.cuh"
#include "allreduce_mc.cuh"
#include "spill_nodes.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
bool g_no_pair = false;               // hcomp_set_cta_pair(0): 1-CTA GEMM tiles only (A/B measurements, tests)
bool g_no_rider_fold = false;        // hcomp_set_rider_fold(0): riders always go through the stand-alone row kernels (A/B, tests)
int g_reserved_sms = 0;               // hcomp_set_reserved_sms(n): SMs the dX GEMM leaves free for a concurrent collective

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define HC_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return fail(HCOMP_E_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define HC_LAUNCH_CHECK(name)                                                                      \
  do {                                                                                             \
    cudaError_t _e = cudaGetLastError();                                                           \
    if (_e != cudaSuccess) return fail(HCOMP_E_CUDA, "launch %s: %s", name, cudaGetErrorString(_e)); \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                            \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct DevInfo {
  int ok = 0, sms = 0, cc_major = 0, cc_minor = 0;
};
int dev_info(DevInfo* out) {
  int dev = 0;
  HC_CUDA(cudaGetDevice(&dev));
  static DevInfo cache[64];
  if (dev < 64 && cache[dev].ok) { *out = cache[dev]; return 0; }
  DevInfo d;
  HC_CUDA(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
  HC_CUDA(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  HC_CUDA(cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (d.cc_major != 10) return fail(HCOMP_E_DEVICE, "device %d is sm_%d%d; this library is sm_100a only", dev, d.cc_major, d.cc_minor);
  d.ok = 1;
  if (dev < 64) cache[dev] = d;
  *out = d;
  return 0;
}

// Internal side branch: independent small kernels of one call run beside each other (fork/join with events on the
// caller's stream: no host synchronisation, capturable into a CUDA graph).
struct SideBranch {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
int side_branch(SideBranch** out, int which = 0) {
  int dev = 0;
  HC_CUDA(cudaGetDevice(&dev));
  static SideBranch cache[64][2];
  if (dev >= 64 || which < 0 || which > 1) return fail(HCOMP_E_ARG, "device index %d / branch %d", dev, which);
  SideBranch& b = cache[dev][which];
  if (b.stream == nullptr) {
    HC_CUDA(cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking));
    HC_CUDA(cudaEventCreateWithFlags(&b.fork, cudaEventDisableTiming));
    HC_CUDA(cudaEventCreateWithFlags(&b.join, cudaEventDisableTiming));
  }
  *out = &b;
  return 0;
}
#define HC_FORK(b, main)                               \
  do {                                                 \
    HC_CUDA(cudaEventRecord((b)->fork, (main)));       \
    HC_CUDA(cudaStreamWaitEvent((b)->stream, (b)->fork, 0)); \
  } while (0)
#define HC_JOIN(b, main)                               \
  do {                                                 \
    HC_CUDA(cudaEventRecord((b)->join, (b)->stream));  \
    HC_CUDA(cudaStreamWaitEvent((main), (b)->join, 0)); \
  } while (0)

// 2D bf16 row-major tensor [outer, inner] (inner contiguous, pitch in elements), 128B-swizzled boxes.
int make_tmap(CUtensorMap* m, const void* base, unsigned long long inner, unsigned long long outer,
              unsigned long long pitch_elems, unsigned box_inner, unsigned box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(HCOMP_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(HCOMP_E_ARG, "operand base not 16-byte aligned");
  if ((pitch_elems * 2) % 16 != 0) return fail(HCOMP_E_ARG, "row pitch %llu bytes is not a multiple of 16", pitch_elems * 2);
  if (box_inner * 2 != 128 || box_outer > 256) return fail(HCOMP_E_ARG, "bad TMA box");
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(HCOMP_E_CUDA, "cuTensorMapEncodeTiled failed with %d (inner %llu outer %llu pitch %llu)", int(r), inner, outer, pitch_elems);
  return 0;
}

// 4D bf16 store map over the compact dZ matrix [n_imgs * HW, pitch]: {w columns of a tile, n_tiles tiles at `w`-column
// pitch, location, image}; boxes of [64 cols x 1 tile x 32 locations x 1 image].
int make_tmap_dz4(CUtensorMap* m, const void* base, unsigned w, unsigned n_tiles, unsigned long long HW,
                  unsigned long long n_imgs, unsigned long long pitch_elems) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(HCOMP_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(HCOMP_E_ARG, "dZ tile base not 16-byte aligned");
  if ((pitch_elems * 2) % 16 != 0 || (w * 2) % 16 != 0 || w == 0 || n_tiles == 0 || HW == 0 || n_imgs == 0)
    return fail(HCOMP_E_ARG, "bad compact dZ geometry (w %u, tiles %u, HW %llu, images %llu, pitch %llu)", w, n_tiles, HW, n_imgs, pitch_elems);
  cuuint64_t gdim[4] = {w, n_tiles, HW, n_imgs};
  cuuint64_t gstr[3] = {(cuuint64_t)w * 2, pitch_elems * 2, HW * pitch_elems * 2};
  cuuint32_t box[4] = {64, 1, 32, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(HCOMP_E_CUDA, "cuTensorMapEncodeTiled (dZ) failed with %d (w %u tiles %u HW %llu)", int(r), w, n_tiles, HW);
  return 0;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device): one bit per device index, set once.  A process
// that drives several GPUs (or several host threads racing here) at worst sets the attribute twice.
template <typename Kern>
int ensure_dyn_smem(Kern kern, int smem, std::atomic<unsigned long long>& done) {
  int dev = 0;
  HC_CUDA(cudaGetDevice(&dev));
  const bool tracked = dev >= 0 && dev < 64;
  if (tracked && ((done.load(std::memory_order_acquire) >> dev) & 1ull)) return 0;
  HC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (tracked) done.fetch_or(1ull << dev, std::memory_order_release);
  return 0;
}

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline int cdiv(long long a, long long b) { return int((a + b - 1) / b); }

// Plain or 2-CTA-cluster launch of a persistent kernel; `workers` = CTAs (or clusters) that loop over the items.
template <typename Kern, typename... Args>
int launch_persistent(Kern kern, const char* name, int cluster, int workers, int threads, int smem, cudaStream_t st,
                      Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(workers * cluster));
  cfg.blockDim = dim3(unsigned(threads));
  cfg.dynamicSmemBytes = size_t(smem);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = unsigned(cluster);
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = cluster > 1 ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
  if (e != cudaSuccess) return fail(HCOMP_E_CUDA, "launch %s: %s", name, cudaGetErrorString(e));
  HC_LAUNCH_CHECK(name);
  return 0;
}

template <int SEG, bool BWD, bool CG2>
int launch_pair(const hc::FeatureMaps& tx, const CUtensorMap& tw,
                const CUtensorMap* td /* [2]: dZ full tiles, partial tile */, const hc::HeadParams& p, int sms, cudaStream_t st) {
  auto kern = hc::head_pair_kernel<SEG, BWD, CG2>;
  constexpr int SMEM = hc::PairMem<BWD, CG2>::SMEM_BYTES;
  static std::atomic<unsigned long long> attr_done{0};
  if (int e = ensure_dyn_smem(kern, SMEM, attr_done)) return e;
  constexpr int CL = CG2 ? 2 : 1;
  const int items = ((p.num_m_tiles + CL - 1) / CL) * p.num_tiles;
  const int slots = sms / CL;
  const int workers = items < slots ? items : slots;
  return launch_persistent(kern, BWD ? "head_pair_kernel<bwd>" : "head_pair_kernel<fwd>", CL, workers,
                           hc::PairCfg<SEG>::THREADS, SMEM, st, tx, tw, td[0], td[1], p);
}

template <bool BWD, bool CG2>
int launch_pair_class(int seg, const hc::FeatureMaps& tx, const CUtensorMap& tw, const CUtensorMap* td,
                      const hc::HeadParams& p, int sms, cudaStream_t st) {
  switch (seg) {
    case 8: return launch_pair<8, BWD, CG2>(tx, tw, td, p, sms, st);
    case 16: return launch_pair<16, BWD, CG2>(tx, tw, td, p, sms, st);
    case 20: return launch_pair<20, BWD, CG2>(tx, tw, td, p, sms, st);
    case 32: return launch_pair<32, BWD, CG2>(tx, tw, td, p, sms, st);
    case 40: return launch_pair<40, BWD, CG2>(tx, tw, td, p, sms, st);
    case 64: return launch_pair<64, BWD, CG2>(tx, tw, td, p, sms, st);
    default: return fail(HCOMP_E_ARG, "unsupported segment class %d (supported: 8,16,20,32,40,64)", seg);
  }
}

// Shared driver of K1 / K5: one launch per segment class (tiles are sorted by class).
template <bool BWD>
int run_pair(const void* x, const void* wp, const int32_t* tiles_host, const int32_t* tiles_dev, int n_tiles, int V,
             int V_first, int HW, int C, int P, int P_pad, int n_nodes, float tau, int precision, hc::HeadParams base,
             const hcomp_spill* spill, bool* riders_folded, cudaStream_t st, bool no_fold = false) {
  *riders_folded = false;
  if (precision != HCOMP_PREC_BF16 && precision != HCOMP_PREC_FP32X3) return fail(HCOMP_E_ARG, "unknown precision %d", precision);
  const int split = (precision == HCOMP_PREC_FP32X3) ? 6 : 1;   // x / wp hold 3 stacked bf16 split planes
  DevInfo di;
  if (int e = dev_info(&di)) return e;
  if (V <= 0 || V_first <= 0 || V_first > V || V - V_first > V_first) return fail(HCOMP_E_ARG, "bad view split V=%d V_first=%d", V, V_first);
  if (HW <= 0) return fail(HCOMP_E_ARG, "HW=%d", HW);
  if (C % 8 != 0 || C <= 0) return fail(HCOMP_E_ARG, "C=%d must be a positive multiple of 8", C);
  if (P_pad != n_tiles * hc::TILE_N) return fail(HCOMP_E_ARG, "P_pad=%d != 128*n_tiles", P_pad);
  if (!(tau > 0.f)) return fail(HCOMP_E_ARG, "softmax tau must be > 0");
  const long long M = (long long)V * HW;
  if (M > 0x7fffffffLL - 2 * hc::TILE_M) return fail(HCOMP_E_ARG, "too many rows");
  const int planes = split > 1 ? 3 : 1;
  if (planes * M > 0x7fffffffLL - 2 * hc::TILE_M) return fail(HCOMP_E_ARG, "too many rows");
  hc::FeatureMaps tx;
  CUtensorMap tw, tw_half;              // tw_half: 64-row boxes for the CTA-pair variant (each CTA keeps half a tile)
  const int cpi = cdiv(HW, 32), rem = HW - 32 * (cpi - 1);
  for (int r = 0; r < 4; ++r) {
    if (int e = make_tmap(&tx.full[r], x, C, planes * M, C, hc::KBLK, 32u * (r + 1))) return e;
    if (int e = make_tmap(&tx.tail[r], x, C, planes * M, C, hc::KBLK, 32u * r + rem)) return e;
  }
  if (int e = make_tmap(&tw, wp, C, (unsigned long long)planes * P_pad, C, hc::KBLK, hc::TILE_N)) return e;
  if (int e = make_tmap(&tw_half, wp, C, (unsigned long long)planes * P_pad, C, hc::KBLK, hc::TILE_N / 2)) return e;
  hc::HeadParams p = base;
  p.M = int(M);
  p.halfM = V_first * HW;
  p.rowsB = int(M) - p.halfM;
  CUtensorMap td[2];                    // backward: dZ leaves through TMA stores (per class: full tiles / partial tile)
  memset(td, 0, sizeof(td));
  p.HW = HW; p.C = C; p.P = P; p.P_pad = P_pad;
  p.split_terms = split;
  p.num_k_blocks = cdiv(C, hc::KBLK) * split;
  p.cpi = cdiv(HW, 32);
  p.inv_cpi = 1.f / float(p.cpi);
  p.rem = rem;
  p.num_chunks = V_first * p.cpi;
  p.num_m_tiles = cdiv(p.num_chunks, 4);
  p.imgs_second = V - V_first;
  p.imgs_total = V;
  p.n_nodes = n_nodes;
  p.imgs_first = V_first;
  p.scale_log2 = 1.4426950408889634f / tau;
  p.inv_tau = 1.f / tau;
  p.inv_HW = 1.f / float(HW);
  p.tiles = tiles_dev;
  int last_fused_global = -1;             // the last tile with segments owns the pitch padding of the compact dZ axis
  for (int i = 0; i < n_tiles; ++i)
    if (tiles_host[(size_t)i * hc::TILE_INTS + 1] > 0) last_fused_global = i;
  // Riders (narrow spill nodes) are finished in the tail of the LAST fused launch when they all belong to its segment
  // class (layout.py takes them out of the last tile of the last class, so they normally do) -- otherwise by the
  // stand-alone row kernels (run_spill).
  hc::HeadParams riders{};
  if (spill != nullptr && spill->n_spill > 0 && spill->recs_host != nullptr && last_fused_global >= 0 && !g_no_rider_fold && !no_fold) {
    const int last_class = tiles_host[(size_t)last_fused_global * hc::TILE_INTS];
    bool ok = true;
    for (int i = 0; i < spill->n_spill && ok; ++i) {
      const int32_t* r = spill->recs_host + (size_t)i * 8;
      if (r[5] == 0) continue;                 // wide node
      if (r[5] != last_class || riders.n_riders >= hc::MAX_RIDERS || r[1] <= 0 || r[1] > r[5]) { ok = false; break; }
      int* d = riders.rider[riders.n_riders++];
      d[0] = r[0]; d[1] = r[1]; d[2] = r[2]; d[3] = r[3]; d[4] = r[4]; d[5] = r[6];
    }
    if (!ok) riders.n_riders = 0;
  }
  int t = 0;
  while (t < n_tiles) {
    const int seg = tiles_host[(size_t)t * hc::TILE_INTS];
    int e = t;
    while (e < n_tiles && tiles_host[(size_t)e * hc::TILE_INTS] == seg) {
      const int32_t* rec = tiles_host + (size_t)e * hc::TILE_INTS;
      if (rec[1] < 0 || rec[1] * seg > hc::TILE_N || rec[2] % 16 != 0 || rec[2] < rec[1] * seg || rec[2] > hc::TILE_N ||
          rec[4] < 0 || rec[4] % 4 != 0 || (rec[4] > 0 && (rec[5] % 4 != 0 || rec[5] < rec[1] * seg || rec[5] + rec[4] > rec[2])) ||
          (rec[1] == 0 && rec[4] == 0))
        return fail(HCOMP_E_ARG, "malformed tile record %d", e);
      if (!BWD && rec[4] > 0 && (p.zs == nullptr || rec[6] < 0 || rec[6] + rec[4] > p.ldz))
        return fail(HCOMP_E_ARG, "tile %d spills %d columns but the scratch matrix is missing or too narrow", e, rec[4]);
      ++e;
    }
    p.tile_begin = t;
    p.num_tiles = e - t;
    if (BWD) {
      // dedicated spill tiles (no segments) sit at the end of the group: their dZ comes from the row kernels, the
      // recompute GEMM skips them
      while (p.num_tiles > 0 && tiles_host[(size_t)(t + p.num_tiles - 1) * hc::TILE_INTS + 1] == 0) --p.num_tiles;
      if (p.num_tiles == 0) { t = e; continue; }
    }
    const int e_fused = t + p.num_tiles;
    if (BWD) {
      // compact dZ columns of this class: full tiles (all segments used) are contiguous at one pitch, then at most one
      // partial tile; widths are the used columns rounded up to 8 (layout.py)
      const int per_tile = hc::TILE_N / seg;
      const int32_t* last = tiles_host + (size_t)(e_fused - 1) * hc::TILE_INTS;
      p.w_full = ((per_tile * seg + 7) / 8) * 8;
      // the last tile of the class is "partial" (own store map) if it has fewer segments, or if it is the globally last
      // tile and owns the alignment padding of the dZ pitch (layout.py rounds P_c up to 64 columns when that fits)
      int last_w = ((last[1] * seg + 7) / 8) * 8;
      if (last_fused_global == e_fused - 1) last_w = p.P_c - last[3];
      const bool has_partial = last_w != p.w_full;
      p.n_full_tiles = p.num_tiles - (has_partial ? 1 : 0);
      p.w_partial = has_partial ? last_w : 0;
      if (has_partial && (last_w <= 0 || last_w > hc::TILE_N || last_w % 8 != 0 || last_w < last[1] * seg))
        return fail(HCOMP_E_ARG, "last tile of class %d: compact width %d", seg, last_w);
      const int c_full = tiles_host[(size_t)t * hc::TILE_INTS + 3];
      for (int i = t; i < e_fused; ++i) {
        const int want = (i - t) < p.n_full_tiles ? c_full + (i - t) * p.w_full : c_full + p.n_full_tiles * p.w_full;
        if (tiles_host[(size_t)i * hc::TILE_INTS + 3] != want || want % 8 != 0)
          return fail(HCOMP_E_ARG, "tile %d: compact dZ column %d, expected %d", i, tiles_host[(size_t)i * hc::TILE_INTS + 3], want);
      }
      if (c_full + p.n_full_tiles * p.w_full + p.w_partial > p.P_c) return fail(HCOMP_E_ARG, "compact dZ columns exceed P_c=%d", p.P_c);
      if (p.n_full_tiles > 0)
        if (int er = make_tmap_dz4(&td[0], p.dz + c_full, p.w_full, p.n_full_tiles, HW, V, p.P_c)) return er;
      if (has_partial) {
        const int c_part = c_full + p.n_full_tiles * p.w_full;
        if (int er = make_tmap_dz4(&td[1], p.dz + c_part, p.w_partial, 1, HW, V, p.P_c)) return er;
      }
      if (p.n_full_tiles == 0) td[0] = td[1];
      if (!has_partial) td[1] = td[0];
    }
    p.n_riders = 0;
    if (riders.n_riders > 0 && last_fused_global >= t && last_fused_global < e) {
      for (int i = 0; i < riders.n_riders; ++i) {
        const int* r = riders.rider[i];
        if (r[3] < 0 || r[3] + r[1] > p.ldz || r[0] < 0 || r[0] >= n_nodes || r[2] < 0 || r[2] + r[1] > P || p.zs == nullptr)
          return fail(HCOMP_E_ARG, "malformed rider record %d", i);
        if (BWD && (r[4] < 0 || r[4] % 8 != 0 || r[5] < r[1] || r[5] % 8 != 0 || r[4] + r[5] > p.P_c))
          return fail(HCOMP_E_ARG, "rider record %d: dZ columns [%d, +%d) outside P_c=%d", i, r[4], r[5], p.P_c);
      }
      p.n_riders = riders.n_riders;
      memcpy(p.rider, riders.rider, sizeof(p.rider));
      *riders_folded = true;
    }
    // CTA pairs (cta_group::2) whenever there are at least two pair tiles along M
    const bool pair = !g_no_pair && p.num_m_tiles >= 2;
    if (int err = pair ? launch_pair_class<BWD, true>(seg, tx, tw_half, td, p, di.sms, st)
                       : launch_pair_class<BWD, false>(seg, tx, tw, td, p, di.sms, st))
      return err;
    t = e;
  }
  return 0;
}

template <bool A_MN, bool B_MN, int OUT>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const hc::GemmParams& p, int sms,
                cudaStream_t st) {
  auto kern = hc::gemm_tc_kernel<A_MN, B_MN, OUT>;
  constexpr int SMEM = hc::GemmCfg<OUT>::SMEM_BYTES;
  static std::atomic<unsigned long long> attr_done{0};
  if (int e = ensure_dyn_smem(kern, SMEM, attr_done)) return e;
  const int items = p.num_m_tiles * p.num_n_tiles * p.splits;
  const int workers = items < sms ? items : sms;
  return launch_persistent(kern, "gemm_tc_kernel", 1, workers, hc::G_THREADS, SMEM, st, ta, tb, to, p);
}

template <bool A_MN, bool B_MN, int OUT>
int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const hc::GemmParams& p, int sms,
                 cudaStream_t st) {
  auto kern = hc::gemm2_tc_kernel<A_MN, B_MN, OUT>;
  constexpr int SMEM = hc::Gemm2Cfg<OUT>::SMEM_BYTES;
  static std::atomic<unsigned long long> attr_done{0};
  if (int e = ensure_dyn_smem(kern, SMEM, attr_done)) return e;
  const int m2 = (p.M + 2 * hc::G_BM - 1) / (2 * hc::G_BM);
  const int items = m2 * p.num_n_tiles * p.splits;
  const int slots = sms / 2;
  const int workers = items < slots ? items : slots;
  return launch_persistent(kern, "gemm2_tc_kernel", 2, workers, hc::G_THREADS, SMEM, st, ta, tb, to, p);
}

// D[M,N] = A[M,K] * B[K,N].  a_mn: A stored [K,M] (M contiguous) else [M,K]; b_mn: B stored [K,N] (N contiguous)
// else [N,K].  splits <= 0 picks a split-K factor that fills the GPU (only meaningful for OUT_RED_F32).
int run_gemm(const void* a, const void* b, long long M, int N, long long K, bool a_mn, bool b_mn, int out_mode, int splits,
             void* out, long long ldo, const int32_t* row_map, cudaStream_t st, int reserve_sms = 0,
             const uint8_t* kact = nullptr, int kact_ld = 0) {
  DevInfo di;
  if (int e = dev_info(&di)) return e;
  if (M <= 0 || N <= 0 || K <= 0) return fail(HCOMP_E_ARG, "empty GEMM");
  if (M > 0x7fffffffLL || K > 0x7fffffffLL) return fail(HCOMP_E_ARG, "GEMM dimension overflow");
  if ((out_mode == hc::OUT_BF16 && (N % 8 || ldo % 8)) || (out_mode != hc::OUT_BF16 && (N % 4 || ldo % 4)))
    return fail(HCOMP_E_ARG, "N=%d / ldo=%lld alignment", N, ldo);
  // CTA-pair (cta_group::2, 256 x 256 tiles) variant whenever there are at least two 128-row tiles
  const bool pair = !g_no_pair && M > hc::G_BM;
  CUtensorMap ta, tb, to;
  if (out_mode == hc::OUT_BF16) {          // bf16 output leaves through TMA stores of [128 rows x 64 cols] boxes
    if (int e = make_tmap(&to, out, N, M, ldo, 64, hc::G_BM)) return e;
  } else {
    memset(&to, 0, sizeof(to));
  }
  if (a_mn) { if (int e = make_tmap(&ta, a, M, K, M, 64, 64)) return e; }
  else      { if (int e = make_tmap(&ta, a, K, M, K, hc::G_BK, hc::G_BM)) return e; }
  if (b_mn) { if (int e = make_tmap(&tb, b, N, K, N, 64, 64)) return e; }
  else      { if (int e = make_tmap(&tb, b, K, N, K, hc::G_BK, pair ? hc::G_BN / 2 : hc::G_BN)) return e; }
  hc::GemmParams p{};
  p.M = int(M); p.N = N; p.K = int(K);
  p.num_m_tiles = cdiv(M, hc::G_BM);
  p.num_n_tiles = cdiv(N, hc::G_BN);
  p.num_k_blocks = cdiv(K, hc::G_BK);
  const int tiles_mn = (pair ? cdiv(M, 2 * hc::G_BM) : p.num_m_tiles) * p.num_n_tiles;
  if (out_mode != hc::OUT_RED_F32) splits = 1;
  int sms = di.sms - reserve_sms;       // persistent grid; a reserve leaves whole SMs to kernels of other streams
  if (sms < 2) sms = 2;
  if (splits <= 0) {
    // split-K factor: minimise rounds x (k-blocks per split + the epilogue's red.add pass, ~8 k-blocks worth), where
    // rounds = ceil(items / workers).  (items / workers rounded down -- the first version -- left cub190's dW with 45 items
    // on 74 CTA pairs in ONE round of 676 k-blocks: 230 us; 8 splits = 5 rounds of 85: 0.63 of that.)
    const int workers = pair ? sms / 2 : sms;
    long long best_cost = -1;
    for (int sp = 1; sp <= 32 && sp <= p.num_k_blocks; ++sp) {
      const int kps = cdiv(p.num_k_blocks, sp);
      if (sp > 1 && kps < 8) break;
      const long long rounds = cdiv((long long)tiles_mn * cdiv(p.num_k_blocks, kps), workers);
      const long long cost = rounds * (kps + 8);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; splits = sp; }
    }
  }
  if (splits < 1) splits = 1;
  if (splits > p.num_k_blocks) splits = p.num_k_blocks;
  p.k_blocks_per_split = cdiv(p.num_k_blocks, splits);
  p.splits = cdiv(p.num_k_blocks, p.k_blocks_per_split);
  p.out = out; p.ldo = ldo; p.row_map = row_map;
  if (kact != nullptr && pair) {          // block-sparse A (only the CTA-pair kernels read the table)
    if (kact_ld % 8 != 0 || kact_ld < p.num_k_blocks || (reinterpret_cast<uintptr_t>(kact) & 7) != 0)
      return fail(HCOMP_E_ARG, "k-block activity table: ld=%d must be a multiple of 8 and >= %d, 8-byte aligned", kact_ld, p.num_k_blocks);
    p.kact = kact; p.kact_ld = kact_ld;
  }
#define HC_GEMM_CASE(AM, BM, OM)                                         \
  if (a_mn == AM && b_mn == BM && out_mode == OM)                        \
    return pair ? launch_gemm2<AM, BM, OM>(ta, tb, to, p, sms, st)       \
                : launch_gemm<AM, BM, OM>(ta, tb, to, p, sms, st);
  HC_GEMM_CASE(false, true, hc::OUT_BF16)      // dX
  HC_GEMM_CASE(true, true, hc::OUT_RED_F32)    // dW
  HC_GEMM_CASE(false, false, hc::OUT_F32)      // self-test: plain K-major GEMM
  HC_GEMM_CASE(false, true, hc::OUT_F32)
  HC_GEMM_CASE(true, true, hc::OUT_F32)
  HC_GEMM_CASE(true, false, hc::OUT_F32)
#undef HC_GEMM_CASE
  return fail(HCOMP_E_ARG, "GEMM variant (a_mn=%d b_mn=%d out=%d) not instantiated", int(a_mn), int(b_mn), out_mode);
}

// Row kernels of the spill nodes (csrc/spill_nodes.cuh): forward = softmax / max-pool / align from the raw logits K1 wrote
// to zs; backward = their dZ columns (no GEMM recompute).
template <bool BWD>
int run_spill(const hcomp_spill* sp, const hc::SpillParams& base, int V, bool riders_folded, cudaStream_t st) {
  if (sp == nullptr || sp->n_spill <= 0) return 0;
  if (sp->recs_host == nullptr || sp->zs == nullptr || sp->ldz <= 0 || sp->ldz % 4 != 0)
    return fail(HCOMP_E_ARG, "spill nodes: records / scratch matrix missing (ldz=%d)", sp->ldz);
  int wide_idx = 0;
  for (int i = 0; i < sp->n_spill; ++i) {
    const int32_t* r = sp->recs_host + (size_t)i * 8;
    hc::SpillParams p = base;
    p.zs = sp->zs; p.ldz = sp->ldz;
    p.node = r[0]; p.P_n = r[1]; p.poff = r[2]; p.zoff = r[3]; p.dz_col = r[4]; p.dz_width = r[6];
    const int cls = r[5];
    if (p.P_n <= 0 || p.zoff < 0 || p.zoff + p.P_n > p.ldz || p.node < 0 || p.node >= p.n_nodes || p.poff < 0 || p.poff + p.P_n > p.P)
      return fail(HCOMP_E_ARG, "malformed spill record %d", i);
    if (BWD && (p.dz_col < 0 || p.dz_col % 8 != 0 || p.dz_width < p.P_n || p.dz_width % 8 != 0 || p.dz_col + p.dz_width > p.P_c))
      return fail(HCOMP_E_ARG, "spill record %d: dZ columns [%d, +%d) outside P_c=%d", i, p.dz_col, p.dz_width, p.P_c);
    if (cls == 0) {                       // wide node
      if (sp->stats == nullptr) return fail(HCOMP_E_ARG, "spill nodes: statistics workspace missing");
      p.stats = sp->stats + (size_t)wide_idx * 2 * (size_t)p.M;
      ++wide_idx;
      if (!BWD) {
        hc::spill_wide_stats_kernel<<<cdiv(p.halfM, 8), 256, 0, st>>>(p);
        HC_LAUNCH_CHECK("spill_wide_stats");
        hc::spill_wide_pool_kernel<<<dim3(cdiv(p.P_n, 128), V), 128, 0, st>>>(p);
        HC_LAUNCH_CHECK("spill_wide_pool");
      } else {
        hc::spill_wide_bwd_kernel<<<cdiv(p.halfM, 8), 256, 0, st>>>(p);
        HC_LAUNCH_CHECK("spill_wide_bwd");
      }
      continue;
    }
    if (riders_folded) continue;          // finished in the tail of the fused kernel
    if (p.P_n > cls) return fail(HCOMP_E_ARG, "spill record %d: %d prototypes in class %d", i, p.P_n, cls);
#define HC_SPILL_CASE(SEG)                                                                    \
    case SEG:                                                                                 \
      if (!BWD) hc::spill_narrow_fwd_kernel<SEG><<<cdiv(cdiv(p.halfM, 32), 8), 256, 0, st>>>(p); \
      else hc::spill_narrow_bwd_kernel<SEG><<<cdiv(p.halfM, 256), 256, 0, st>>>(p);           \
      break;
    switch (cls) {
      HC_SPILL_CASE(8) HC_SPILL_CASE(16) HC_SPILL_CASE(20) HC_SPILL_CASE(32) HC_SPILL_CASE(40) HC_SPILL_CASE(64)
      default: return fail(HCOMP_E_ARG, "spill record %d: unsupported segment class %d", i, cls);
    }
#undef HC_SPILL_CASE
    HC_LAUNCH_CHECK(BWD ? "spill_narrow_bwd" : "spill_narrow_fwd");
  }
  return 0;
}

hc::DzBlockTables dz_tables(const hcomp_dz_blocks* blk, int HW) {
  hc::DzBlockTables bt{};
  if (blk != nullptr && blk->t1 != nullptr) {
    bt.t1 = blk->t1; bt.ld1 = blk->ld1; bt.t2 = blk->t2; bt.ld2 = blk->ld2; bt.pcol = blk->pcol;
    if (blk->iact != nullptr && blk->tile_of_node != nullptr) {
      bt.iact = blk->iact; bt.iact_pitch = blk->iact_pitch; bt.tile_of_node = blk->tile_of_node; bt.cpi = (HW + 31) / 32;
    }
  }
  return bt;
}

hc::SpillParams spill_base(int V, int V_first, int HW, int P, int n_nodes, float tau) {
  hc::SpillParams q{};
  q.M = V * HW; q.halfM = V_first * HW; q.rowsB = q.M - q.halfM; q.HW = HW; q.P = P; q.n_nodes = n_nodes;
  q.imgs_first = V_first;
  q.scale_log2 = 1.4426950408889634f / tau; q.inv_tau = 1.f / tau; q.inv_HW = 1.f / float(HW);
  return q;
}

inline int blocks(long long n, int bs) { return int((n + bs - 1) / bs); }

}  // namespace

extern "C" {

int hcomp_abi_version(void) { return HCOMP_ABI_VERSION; }
int hcomp_set_cta_pair(int on) {
  const int prev = g_no_pair ? 0 : 1;
  g_no_pair = (on == 0);
  return prev;
}
int hcomp_set_rider_fold(int on) {
  const int prev = g_no_rider_fold ? 0 : 1;
  g_no_rider_fold = (on == 0);
  return prev;
}
int hcomp_set_reserved_sms(int n) {
  const int prev = g_reserved_sms;
  g_reserved_sms = n < 0 ? 0 : n;
  return prev;
}
int hcomp_split3_f32(const float* src, void* dst_bf16_3planes, long long n, void* stream) {
  int grid = blocks(n, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  hc::split3_f32_kernel<<<grid, 256, 0, S(stream)>>>(src, reinterpret_cast<__nv_bfloat16*>(dst_bf16_3planes), n);
  HC_LAUNCH_CHECK("split3_f32");
  return 0;
}
int hcomp_pack_weights_split3(const float* w_flat, const int32_t* row_map, int P_pad, int C, void* wp3_bf16, void* stream) {
  const long long n = (long long)P_pad * C;
  hc::pack_weights_split3_kernel<<<blocks(n, 256), 256, 0, S(stream)>>>(w_flat, row_map, P_pad, C,
                                                                       reinterpret_cast<__nv_bfloat16*>(wp3_bf16));
  HC_LAUNCH_CHECK("pack_weights_split3");
  return 0;
}

const char* hcomp_last_error(void) { return g_err; }
long long hcomp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int hcomp_num_sms(void) {
  DevInfo di;
  if (int e = dev_info(&di)) return e;
  return di.sms;
}
int hcomp_init(void) {
  DevInfo di;
  if (int e = dev_info(&di)) return e;
  SideBranch* b = nullptr;
  for (int which = 0; which < 2; ++which)
    if (int e = side_branch(&b, which)) return e;
  return 0;
}

int hcomp_pack_weights(const float* w_flat, const int32_t* row_map, int P_pad, int C, void* wp_bf16, void* stream) {
  if (C % 8) return fail(HCOMP_E_ARG, "C must be a multiple of 8");
  const long long n = (long long)P_pad * (C / 8);
  hc::pack_weights_kernel<<<blocks(n, 256), 256, 0, S(stream)>>>(w_flat, row_map, P_pad, C,
                                                                  reinterpret_cast<__nv_bfloat16*>(wp_bf16));
  HC_LAUNCH_CHECK("pack_weights");
  return 0;
}

int hcomp_cast_f32_to_bf16(const float* src, void* dst_bf16, long long n, void* stream) {
  if (n % 8) return fail(HCOMP_E_ARG, "n must be a multiple of 8");
  const long long n8 = n / 8;
  int grid = blocks(n8, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  hc::cast_f32_bf16_kernel<<<grid, 256, 0, S(stream)>>>(src, reinterpret_cast<__nv_bfloat16*>(dst_bf16), n8);
  HC_LAUNCH_CHECK("cast_f32_bf16");
  return 0;
}

int hcomp_nchw_to_rows_bf16(const void* src, int src_is_bf16, int V, int C, int HW, void* dst_bf16, void* stream) {
  dim3 grid(cdiv(HW, 32), cdiv(C, 32), V), block(32, 8);
  if (src_is_bf16)
    hc::nchw_to_rows_bf16_kernel<__nv_bfloat16><<<grid, block, 0, S(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(src), reinterpret_cast<__nv_bfloat16*>(dst_bf16), C, HW);
  else
    hc::nchw_to_rows_bf16_kernel<float><<<grid, block, 0, S(stream)>>>(reinterpret_cast<const float*>(src),
                                                                      reinterpret_cast<__nv_bfloat16*>(dst_bf16), C, HW);
  HC_LAUNCH_CHECK("nchw_to_rows_bf16");
  return 0;
}

int hcomp_scale_residual_rows_bf16(const void* y, int y_is_bf16, const void* res, int res_is_bf16, const float* gamma,
                                   const float* keep, int V, int C, int HW, void* out_bf16, void* stream) {
  if (V <= 0 || HW <= 0 || C <= 0 || C % 8 != 0) return fail(HCOMP_E_ARG, "scale_residual: V=%d HW=%d C=%d (C must be a multiple of 8)", V, HW, C);
  if (((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(out_bf16) |
        reinterpret_cast<uintptr_t>(gamma)) & 15) != 0)
    return fail(HCOMP_E_ARG, "scale_residual: operands must be 16-byte aligned");
  const long long n8 = (long long)V * HW * C / 8;
  int grid = blocks(n8, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
#define HC_SR(TY, TR) hc::scale_residual_rows_kernel<TY, TR><<<grid, 256, 0, S(stream)>>>( \
      reinterpret_cast<const TY*>(y), reinterpret_cast<const TR*>(res), gamma, keep, n8, C / 8, HW, out)
  if (y_is_bf16 && res_is_bf16) HC_SR(__nv_bfloat16, __nv_bfloat16);
  else if (y_is_bf16) HC_SR(__nv_bfloat16, float);
  else if (res_is_bf16) HC_SR(float, __nv_bfloat16);
  else HC_SR(float, float);
#undef HC_SR
  HC_LAUNCH_CHECK("scale_residual_rows");
  return 0;
}

int hcomp_label_tables(const long long* ys, const hcomp_tables* t, int V, int V_first, int8_t* tgt, uint8_t* desc,
                       int32_t* n_desc, void* stream) {
  HC_CUDA(cudaMemsetAsync(n_desc, 0, sizeof(int32_t) * t->n_nodes, S(stream)));
  hc::label_tables_kernel<<<blocks((long long)V * t->n_nodes, 256), 256, 0, S(stream)>>>(ys, t->anc, V, V_first, t->n_nodes,
                                                                                        t->n_leaves, tgt, desc, n_desc);
  HC_LAUNCH_CHECK("label_tables");
  return 0;
}

int hcomp_proj_softmax_pool_fwd(const void* x_bf16, const void* wp_bf16, const int32_t* tiles_host,
                                const int32_t* tiles_dev, int n_tiles, int V, int V_first, int HW, int C, int P,
                                int P_pad, int n_nodes, float tau, int precision, int outputs_zeroed,
                                const uint8_t* desc, unsigned long long* pooled_packed, double* align_sum,
                                const hcomp_spill* spill, void* stream) {
  hc::HeadParams p{};
  p.pooled_packed = pooled_packed;
  p.align_sum = align_sum;
  p.desc = (align_sum != nullptr) ? desc : nullptr;
  if (spill != nullptr && spill->n_spill > 0) { p.zs = spill->zs; p.ldz = spill->ldz; }
  if (!(outputs_zeroed & 1)) {
    HC_CUDA(cudaMemsetAsync(pooled_packed, 0, sizeof(unsigned long long) * (size_t)V * P, S(stream)));
    if (align_sum) HC_CUDA(cudaMemsetAsync(align_sum, 0, sizeof(double) * n_nodes, S(stream)));
  }
  // bit 1: the narrow spill nodes (riders) are NOT finished here -- neither in the kernel's tail (no grid barrier, the
  // kernel ends with its item loop) nor by the row kernels: hcomp_pool_classify_fwd does it, given the same `spill`
  const bool defer = (outputs_zeroed & 2) != 0;
  bool folded = false;
  if (int e = run_pair<false>(x_bf16, wp_bf16, tiles_host, tiles_dev, n_tiles, V, V_first, HW, C, P, P_pad, n_nodes, tau,
                              precision, p, spill, &folded, S(stream), defer))
    return e;
  if (!(tau > 0.f)) return fail(HCOMP_E_ARG, "softmax tau must be > 0");
  hc::SpillParams q = spill_base(V, V_first, HW, P, n_nodes, tau);
  q.pooled_packed = pooled_packed; q.align_sum = align_sum; q.desc = p.desc;
  return run_spill<false>(spill, q, V, folded || defer, S(stream));
}

int hcomp_unpack_pool(const unsigned long long* packed, long long n, float thresh, float* pooled, int32_t* argmax,
                      void* stream) {
  hc::unpack_pool_kernel<<<blocks(n, 256), 256, 0, S(stream)>>>(packed, n, thresh, pooled, argmax);
  HC_LAUNCH_CHECK("unpack_pool");
  return 0;
}

int hcomp_align_finalize(const double* align_sum, const int32_t* n_desc, int N, int HW, float* loss, void* stream) {
  hc::align_finalize_kernel<<<blocks(N, 128), 128, 0, S(stream)>>>(align_sum, n_desc, N, HW, loss);
  HC_LAUNCH_CHECK("align_finalize");
  return 0;
}

int hcomp_head_bwd_dz(const void* x_bf16, const void* wp_bf16, const int32_t* tiles_host, const int32_t* tiles_dev,
                      int n_tiles, int V, int V_first, int HW, int C, int P, int P_pad, int P_c, int n_nodes, float tau,
                      int precision, const int32_t* argmax, const float* g_pooled, const float* pooled, float thresh,
                      const uint8_t* desc, const int32_t* n_desc, const float* g_align, void* scat_ws, float* coef_ws,
                      void* dz_bf16, const hcomp_spill* spill, const hcomp_dz_blocks* blk, const int32_t* proto_off,
                      const int32_t* proto_node, void* stream) {
  if (P_c <= 0 || P_c % 8 != 0 || P_c > P_pad) return fail(HCOMP_E_ARG, "P_c=%d must be a positive multiple of 8, <= P_pad", P_c);
  const long long n = (long long)V * P;
  hc::HeadParams p{};
  p.scat = reinterpret_cast<const int2*>(scat_ws);
  p.dz = reinterpret_cast<__nv_bfloat16*>(dz_bf16);
  p.P_c = P_c;
  if (argmax == nullptr) {              // scat_ws / coef_ws were filled by hcomp_head_chain_bwd
    if (g_align != nullptr && desc != nullptr && n_desc != nullptr) p.coef_align = coef_ws;
  } else {                              // scatter table + align coefficients: one launch
    const bool use_align = g_align != nullptr && desc != nullptr && n_desc != nullptr;
    const int nb_scat = blocks(n, 256), nb_coef = use_align ? blocks((long long)V_first * n_nodes, 256) : 0;
    if (blk != nullptr && blk->t1 != nullptr && (proto_off == nullptr || proto_node == nullptr || blk->pcol == nullptr))
      return fail(HCOMP_E_ARG, "dZ block tables need proto_off, proto_node and pcol");
    const hc::DzBlockTables bt = dz_tables(blk, HW);
    hc::bwd_prep_kernel<<<nb_scat + nb_coef, 256, 0, S(stream)>>>(argmax, g_pooled, thresh > 0.f ? pooled : nullptr, thresh, n,
                                                                  reinterpret_cast<int2*>(scat_ws), nb_scat, desc, n_desc,
                                                                  g_align, V_first, n_nodes, HW, coef_ws, P, V, proto_off, proto_node, bt);
    HC_LAUNCH_CHECK("bwd_prep");
    if (use_align) p.coef_align = coef_ws;
  }
  if (spill != nullptr && spill->n_spill > 0) { p.zs = spill->zs; p.ldz = spill->ldz; }
  if (blk != nullptr && blk->t1 != nullptr && blk->iact != nullptr && blk->tile_of_node != nullptr) {
    const int chunks = V_first * ((HW + 31) / 32);
    if (blk->iact_pitch % 8 != 0 || blk->iact_pitch < (chunks + 7) / 8 * 8 + 8 || (reinterpret_cast<uintptr_t>(blk->iact) & 7) != 0)
      return fail(HCOMP_E_ARG, "item activity table: pitch %d for %d chunks (multiple of 8, >= chunks rounded up + 8, 8-byte aligned)",
                  blk->iact_pitch, chunks);
    p.iact = blk->iact; p.iact_pitch = blk->iact_pitch;
    // (the 1-CTA GEMM family reads dZ densely: no unstored tiles when it is selected)
    if (blk->dz_only_read_through_tables && !g_no_pair) { p.gt1 = blk->t1; p.gld1 = blk->ld1; p.gt2 = blk->t2; p.gld2 = blk->ld2; }
  }
  bool folded = false;
  if (int e = run_pair<true>(x_bf16, wp_bf16, tiles_host, tiles_dev, n_tiles, V, V_first, HW, C, P, P_pad, n_nodes, tau,
                             precision, p, spill, &folded, S(stream)))
    return e;
  hc::SpillParams q = spill_base(V, V_first, HW, P, n_nodes, tau);
  q.scat = p.scat; q.coef_align = p.coef_align; q.dz = p.dz; q.P_c = P_c;
  return run_spill<true>(spill, q, V, folded, S(stream));
}

int hcomp_head_bwd_dx(const void* dz_bf16, const void* wpc_bf16, long long rows, int P_c, int C, void* dx_bf16,
                      const hcomp_dz_blocks* blk, void* stream) {
  return run_gemm(dz_bf16, wpc_bf16, rows, C, P_c, false, true, hc::OUT_BF16, 1, dx_bf16, C, nullptr, S(stream),
                  g_reserved_sms, blk ? blk->t1 : nullptr, blk ? blk->ld1 : 0);
}

int hcomp_head_bwd_dw(const void* dz_bf16, const void* x_bf16, const int32_t* row_map_c, long long rows, int P_c, int C,
                      float* dw_flat, const hcomp_dz_blocks* blk, void* stream) {
  return run_gemm(dz_bf16, x_bf16, P_c, C, rows, true, true, hc::OUT_RED_F32, 0, dw_flat, C, row_map_c, S(stream), 0,
                  blk ? blk->t2 : nullptr, blk ? blk->ld2 : 0);
}

int hcomp_classifier_fwd(const float* pooled, const float* wc, const float* bias, const hcomp_tables* t, int V,
                         float* out, void* stream) {
  const long long n = (long long)V * t->n_cols;
  hc::classifier_fwd_kernel<<<blocks(n, 128), 128, 0, S(stream)>>>(pooled, wc, bias, t->col_node, t->proto_off, t->cls_off,
                                                                  t->wc_off, V, t->n_protos, t->n_cols, out);
  HC_LAUNCH_CHECK("classifier_fwd");
  return 0;
}

int hcomp_classifier_bwd(const float* g_out, const float* pooled, const float* wc, const hcomp_tables* t, int V,
                         float* g_pooled, int accumulate, float* g_wc, float* g_bias, void* stream) {
  SideBranch* sb = nullptr;             // the weight / bias gradients are independent of g_pooled: side branch
  cudaStream_t ws = S(stream);
  if (g_pooled && (g_wc || g_bias)) {
    if (int e = side_branch(&sb, 0)) return e;
    HC_FORK(sb, S(stream));
    ws = sb->stream;
  }
  if (g_wc) {
    const long long n = (long long)t->n_welems * 32;
    hc::classifier_bwd_weight_kernel<<<blocks(n, 256), 256, 0, ws>>>(g_out, pooled, wc, t->welem_col, t->welem_proto, V,
                                                                    t->n_protos, t->n_cols, t->n_welems, g_wc, nullptr);
    HC_LAUNCH_CHECK("classifier_bwd_weight");
  }
  if (g_bias) {
    hc::classifier_bwd_bias_kernel<<<blocks(t->n_cols, 128), 128, 0, ws>>>(g_out, V, t->n_cols, g_bias);
    HC_LAUNCH_CHECK("classifier_bwd_bias");
  }
  if (g_pooled) {
    const long long n = (long long)V * t->n_protos;
    hc::classifier_bwd_pooled_kernel<<<blocks(n, 256), 256, 0, S(stream)>>>(g_out, wc, t->proto_node, t->proto_off,
                                                                           t->cls_off, t->wc_off, V, t->n_protos,
                                                                           t->n_cols, g_pooled, accumulate);
    HC_LAUNCH_CHECK("classifier_bwd_pooled");
  }
  if (sb) HC_JOIN(sb, S(stream));
  return 0;
}

namespace {
struct LossWs {            // carve-up of the float workspace shared by hcomp_head_losses_fwd / _bwd
  float *colsum, *tanh_part, *orth_sq, *cls, *E;
};
LossWs loss_ws(float* ws, const hcomp_tables* t) {
  LossWs w;
  w.colsum = ws;
  w.tanh_part = w.colsum + 2 * (size_t)t->n_protos;
  w.orth_sq = w.tanh_part + 2 * (size_t)t->n_nodes;
  w.cls = w.orth_sq + t->n_nodes;
  w.E = w.cls + t->n_nodes;
  return w;
}
}  // namespace

long long hcomp_head_losses_ws_floats(const hcomp_tables* t) {
  return 2LL * t->n_protos + 4LL * t->n_nodes + (long long)t->n_nodes * t->p_max * t->p_max;
}

int hcomp_head_losses_fwd(const float* pooled, const float* out, const float* align, const float* w_flat, const float* wc,
                          const int8_t* tgt, const int32_t* n_desc, const hcomp_tables* t, int V, int V_first, int C,
                          int flags, const float* weights_host, float eps, float multiplier, float* total, float* stats,
                          int32_t* n_correct, float* ws, uint8_t* rel, void* stream) {
  const LossWs w = loss_ws(ws, t);
  const bool do_tanh = flags & HCOMP_LOSS_TANH, do_orth = flags & HCOMP_LOSS_ORTH, do_cls = flags & HCOMP_LOSS_CLASS;
  if ((flags & HCOMP_LOSS_SPARSITY) && !(multiplier > 0.f)) return fail(HCOMP_E_ARG, "class loss: log1p(out**m) needs m > 0 (got %g)", multiplier);
  const float sparsity = (flags & HCOMP_LOSS_SPARSITY) ? multiplier : 0.f;
  SideBranch* sb = nullptr;
  if (do_orth) {        // depends on the weights only: runs on the side branch beside the tanh / class kernels
    if (t->p_max >= C) return fail(HCOMP_E_ARG, "orth loss needs P_n < C (P_max=%d, C=%d)", t->p_max, C);
    if (int e = side_branch(&sb)) return e;
    HC_FORK(sb, S(stream));
    const long long warps = (long long)t->n_nodes * (t->p_max * (t->p_max + 1) / 2);     // upper triangle
    hc::orth_gram_kernel<<<blocks(warps * 32, 256), 256, 0, sb->stream>>>(w_flat, wc, t->proto_off, t->cls_off, t->wc_off,
                                                                         t->n_nodes, C, t->p_max, w.E, rel);
    HC_LAUNCH_CHECK("orth_gram");
    hc::orth_sumsq_kernel<<<blocks((long long)t->n_nodes * 32, 128), 128, 0, sb->stream>>>(w.E, t->proto_off, t->n_nodes,
                                                                                          t->p_max, w.orth_sq);
    HC_LAUNCH_CHECK("orth_sumsq");
  }
  SideBranch* sb_t = nullptr;
  if (do_tanh) {        // second side branch: the class kernel below stays on the caller's stream
    if (t->p_max > 64 * 1024) return fail(HCOMP_E_ARG, "P_max too large");
    if (int e = side_branch(&sb_t, 1)) return e;
    HC_FORK(sb_t, S(stream));
    hc::tanh_loss_fwd_kernel<<<dim3(t->n_nodes, 2), 256, 0, sb_t->stream>>>(pooled, tgt, t->proto_off, n_desc, V, V_first,
                                                                           t->n_nodes, t->n_protos, eps, w.tanh_part, w.colsum);
    HC_LAUNCH_CHECK("tanh_loss_fwd");
  }
  // class kernel also produces the per-node accuracy counters, so it always runs
  hc::class_loss_fwd_kernel<<<t->n_nodes, 128, 0, S(stream)>>>(out, tgt, t->child_w, t->cls_off, n_desc, V, t->n_nodes,
                                                              t->n_cols, sparsity, w.cls, n_correct);
  HC_LAUNCH_CHECK("class_loss_fwd");
  if (sb) HC_JOIN(sb, S(stream));
  if (sb_t) HC_JOIN(sb_t, S(stream));
  hc::LossWeights lw;
  for (int i = 0; i < 4; ++i) lw.w[i] = weights_host[i];
  hc::loss_combine_kernel<<<1, 256, 0, S(stream)>>>(align, do_tanh ? w.tanh_part : nullptr, do_orth ? w.orth_sq : nullptr,
                                                   do_cls ? w.cls : nullptr, n_desc, t->n_nodes, lw, stats, total);
  HC_LAUNCH_CHECK("loss_combine");
  return 0;
}

int hcomp_head_losses_bwd(const float* g_total, const float* out, const float* w_flat, const int8_t* tgt,
                          const int32_t* n_desc, const float* stats, const hcomp_tables* t, int V, int V_first, int C,
                          int flags, const float* weights_host, float eps, float multiplier, const float* ws,
                          const uint8_t* rel, float* gvec, float* g_pooled, float* g_out, float* g_w, void* stream) {
  const LossWs w = loss_ws(const_cast<float*>(ws), t);
  const int N = t->n_nodes;
  if ((flags & HCOMP_LOSS_SPARSITY) && !(multiplier > 0.f)) return fail(HCOMP_E_ARG, "class loss: log1p(out**m) needs m > 0 (got %g)", multiplier);
  const float sparsity = (flags & HCOMP_LOSS_SPARSITY) ? multiplier : 0.f;
  hc::LossWeights lw;
  for (int i = 0; i < 4; ++i) lw.w[i] = weights_host[i];
  hc::loss_grads_kernel<<<blocks(4 * N, 128), 128, 0, S(stream)>>>(g_total, N, lw, gvec);   // gvec[0..N) is g_align
  HC_LAUNCH_CHECK("loss_grads");
  SideBranch* sb_bwd = nullptr;
  if (g_w) {
    if (flags & HCOMP_LOSS_ORTH) {      // weights-only term: beside the tanh / class backward kernels
      if (int e = side_branch(&sb_bwd)) return e;
      HC_FORK(sb_bwd, S(stream));
      hc::orth_bwd_kernel<<<dim3(t->n_protos, (C + 255) / 256), 256, 0, sb_bwd->stream>>>(
          w_flat, t->proto_node, t->proto_off, C, t->p_max, stats + 2 * N, w.E, rel, gvec + 2 * N, g_w);
      HC_LAUNCH_CHECK("orth_bwd");
    } else {
      HC_CUDA(cudaMemsetAsync(g_w, 0, sizeof(float) * (size_t)t->n_protos * C, S(stream)));
    }
  }
  SideBranch* sb_tb = nullptr;
  if (g_pooled) {
    if (flags & HCOMP_LOSS_TANH) {
      const long long n = (long long)V * t->n_protos;
      if (int e = side_branch(&sb_tb, 1)) return e;
      HC_FORK(sb_tb, S(stream));
      hc::tanh_loss_bwd_kernel<<<blocks(n, 256), 256, 0, sb_tb->stream>>>(w.colsum, tgt, t->proto_node, t->proto_off, gvec + N,
                                                                         V, V_first, N, t->n_protos, eps, g_pooled, 0);
      HC_LAUNCH_CHECK("tanh_loss_bwd");
    } else {
      HC_CUDA(cudaMemsetAsync(g_pooled, 0, sizeof(float) * (size_t)V * t->n_protos, S(stream)));
    }
  }
  if (g_out) {
    if (flags & HCOMP_LOSS_CLASS) {
      const long long n = (long long)V * t->n_cols;
      hc::class_loss_bwd_kernel<<<blocks(n, 128), 128, 0, S(stream)>>>(out, tgt, t->child_w, t->col_node, t->cls_off, n_desc,
                                                                      gvec + 3 * N, V, N, t->n_cols, sparsity, g_out);
      HC_LAUNCH_CHECK("class_loss_bwd");
    } else {
      HC_CUDA(cudaMemsetAsync(g_out, 0, sizeof(float) * (size_t)V * t->n_cols, S(stream)));
    }
  }
  if (sb_bwd) HC_JOIN(sb_bwd, S(stream));
  if (sb_tb) HC_JOIN(sb_tb, S(stream));
  return 0;
}

// ---------------------------------------------------------------- fused chains (ABI v6)
int hcomp_head_prologue(const float* w_flat, const int32_t* row_map, int rows, int C, void* wp_bf16,
                        unsigned long long* packed, long long n_packed, double* align_sum, int n_align, const long long* ys,
                        const hcomp_tables* t, int V, int V_first, int8_t* tgt, uint8_t* desc, int32_t* n_desc,
                        void* zero_extra, long long zero_extra_bytes, void* stream) {
  hc::PrologueParams q{};
  if (zero_extra != nullptr && zero_extra_bytes > 0) {
    if (zero_extra_bytes % 16 != 0 || (reinterpret_cast<uintptr_t>(zero_extra) & 15) != 0)
      return fail(HCOMP_E_ARG, "prologue: the extra buffer to clear must be 16-byte aligned and a multiple of 16 bytes");
    q.zero16 = reinterpret_cast<uint4*>(zero_extra); q.n_zero16 = zero_extra_bytes / 16;
  }
  if (rows > 0) {
    if (C % 8 != 0) return fail(HCOMP_E_ARG, "prologue: C=%d must be a multiple of 8", C);
    q.w = w_flat; q.row_map = row_map; q.rows = rows; q.C = C; q.wp = reinterpret_cast<__nv_bfloat16*>(wp_bf16);
    q.nb_pack = blocks((long long)rows * (C >> 3), 256);
  }
  if (packed != nullptr || align_sum != nullptr || q.n_zero16 > 0) {
    q.packed = packed; q.n_packed = packed ? n_packed : 0; q.align_sum = align_sum; q.n_align = n_align;
    q.nb_zero = blocks((q.n_packed + 1) / 2, 256);
    if (q.nb_zero == 0) q.nb_zero = 1;
  }
  if (ys != nullptr) {
    q.ys = ys; q.anc = t->anc; q.V = V; q.V_first = V_first; q.N = t->n_nodes; q.L = t->n_leaves;
    q.tgt = tgt; q.desc = desc; q.n_desc = n_desc;
    q.nb_tgt = blocks((long long)V * t->n_nodes, 256);
    q.nb_cnt = blocks(t->n_nodes, 8);
  }
  const int grid = q.nb_pack + q.nb_zero + q.nb_tgt + q.nb_cnt;
  if (grid == 0) return 0;
  hc::head_prologue_kernel<<<grid, 256, 0, S(stream)>>>(q);
  HC_LAUNCH_CHECK("head_prologue");
  return 0;
}

int hcomp_pool_classify_fwd(const unsigned long long* packed, double* align_sum, const int32_t* n_desc, const float* wc,
                            const float* bias, const hcomp_tables* t, int V, int HW, float thresh, float* pooled,
                            int32_t* argmax, float* align, float* out, const hcomp_spill* deferred, int V_first, float tau,
                            const uint8_t* desc, unsigned int* counter, void* stream) {
  hc::PoolClassifyParams q{};
  q.packed = packed; q.align_sum = align_sum; q.n_desc = n_desc; q.wc = wc; q.bias = bias;
  q.col_node = t->col_node; q.proto_off = t->proto_off; q.cls_off = t->cls_off; q.wc_off = t->wc_off;
  q.V = V; q.P = t->n_protos; q.K = t->n_cols; q.N = t->n_nodes; q.HW = HW; q.thresh = thresh;
  q.pooled = pooled; q.argmax = argmax; q.out = (out != nullptr && wc != nullptr) ? out : nullptr;
  q.align = (align != nullptr && align_sum != nullptr && n_desc != nullptr) ? align : nullptr;
  int nb_rider = 0;
  const bool have_riders = deferred != nullptr && deferred->n_spill > 0;
  const int BS = have_riders ? 512 : 256;               // block size of the two kernel variants
  q.nb_unpack = blocks((long long)V * t->n_protos, BS);
  q.nb_cls = q.out ? blocks((long long)V * t->n_cols, BS) : 0;
  q.nb_align = q.align ? blocks(t->n_nodes, BS) : 0;
  if (have_riders) {
    if (deferred->recs_host == nullptr || deferred->zs == nullptr || deferred->ldz <= 0 || deferred->ldz % 4 != 0)
      return fail(HCOMP_E_ARG, "deferred riders: records / scratch matrix missing (ldz=%d)", deferred->ldz);
    if (!(tau > 0.f) || V_first <= 0 || V_first > V || counter == nullptr)
      return fail(HCOMP_E_ARG, "deferred riders: need tau > 0, 0 < V_first <= V and a counter block");
    const hc::SpillParams base = spill_base(V, V_first, HW, t->n_protos, t->n_nodes, tau);
    for (int i = 0; i < deferred->n_spill; ++i) {
      const int32_t* r = deferred->recs_host + (size_t)i * 8;
      if (r[5] == 0) return fail(HCOMP_E_ARG, "deferred riders: record %d is a wide node (finished by the forward call)", i);
      if (q.n_riders >= 4) return fail(HCOMP_E_ARG, "deferred riders: at most 4");
      hc::SpillParams& p = q.rider[q.n_riders];
      p = base;
      p.zs = deferred->zs; p.ldz = deferred->ldz;
      p.node = r[0]; p.P_n = r[1]; p.poff = r[2]; p.zoff = r[3];
      p.pooled_packed = const_cast<unsigned long long*>(packed);
      p.align_sum = q.align ? align_sum : nullptr;
      p.desc = q.align ? desc : nullptr;
      if (p.P_n <= 0 || p.P_n > 64 || p.zoff < 0 || p.zoff + p.P_n > p.ldz || p.node < 0 || p.node >= p.n_nodes || p.poff < 0 ||
          p.poff + p.P_n > p.P)
        return fail(HCOMP_E_ARG, "malformed deferred rider record %d", i);
      ++q.n_riders;
    }
    q.align_sum_rw = align_sum; q.counter = counter;
    nb_rider = q.n_riders * V_first;
  }
  if (nb_rider > 0) {
    int pmax = 1;
    for (int i = 0; i < q.n_riders; ++i) pmax = q.rider[i].P_n > pmax ? q.rider[i].P_n : pmax;
    // rider_finish_block's carve-up: the pair's logits [2][HW][P_n | 1], row statistics [4][HW], slice results, scratch
    const size_t smem = sizeof(float) * ((size_t)2 * HW * (pmax | 1) + (size_t)4 * HW + 512 + 512 + 128 + 32);
    constexpr int SMEM_CAP = 200 * 1024;
    if (smem > SMEM_CAP)
      return fail(HCOMP_E_ARG, "deferred riders: HW=%d x P_n=%d does not fit the finish kernel's shared memory (clear bit 1)", HW, pmax);
    static std::atomic<unsigned long long> attr_done{0};
    if (int e = ensure_dyn_smem(hc::pool_classify_fwd_kernel<true>, SMEM_CAP, attr_done)) return e;
    hc::pool_classify_fwd_kernel<true><<<q.nb_unpack + q.nb_cls + q.nb_align + nb_rider, 512, smem, S(stream)>>>(q);
  } else {
    hc::pool_classify_fwd_kernel<false><<<q.nb_unpack + q.nb_cls + q.nb_align, 256, 0, S(stream)>>>(q);
  }
  HC_LAUNCH_CHECK("pool_classify_fwd");
  return 0;
}

long long hcomp_head_chain_ws_floats(const hcomp_tables* t, int V) {
  return hcomp_head_losses_ws_floats(t) + (long long)V * t->n_nodes;
}

int hcomp_orth_gram(const float* w_flat, const float* wc, const hcomp_tables* t, int C, float* ws, uint8_t* rel, void* stream) {
  const LossWs w = loss_ws(ws, t);
  if (t->p_max >= C) return fail(HCOMP_E_ARG, "orth loss needs P_n < C (P_max=%d, C=%d)", t->p_max, C);
  const long long warps = (long long)t->n_nodes * (t->p_max * (t->p_max + 1) / 2);     // upper triangle
  hc::orth_gram_kernel<<<blocks(warps * 32, 256), 256, 0, S(stream)>>>(w_flat, wc, t->proto_off, t->cls_off, t->wc_off,
                                                                      t->n_nodes, C, t->p_max, w.E, rel);
  HC_LAUNCH_CHECK("orth_gram");
  return 0;
}

int hcomp_head_chain_fwd(const float* pooled, const float* out, const float* align, const float* w_flat, const float* wc,
                         const int8_t* tgt, const int32_t* n_desc, const hcomp_tables* t, int V, int V_first, int C,
                         int flags, const float* weights_host, float eps, float multiplier, float* total, float* stats,
                         int32_t* n_correct, float* ws, uint8_t* rel, unsigned int* counter, void* stream) {
  const LossWs w = loss_ws(ws, t);
  const bool do_tanh = flags & HCOMP_LOSS_TANH, do_orth = flags & HCOMP_LOSS_ORTH, do_cls = flags & HCOMP_LOSS_CLASS;
  if ((flags & HCOMP_LOSS_SPARSITY) && !(multiplier > 0.f)) return fail(HCOMP_E_ARG, "class loss: log1p(out**m) needs m > 0 (got %g)", multiplier);
  if (do_orth && !(flags & HCOMP_LOSS_ORTH_READY))
    if (int e = hcomp_orth_gram(w_flat, wc, t, C, ws, rel, stream)) return e;
  hc::ChainFwdParams q{};
  q.pooled = pooled; q.out = out; q.align = align; q.orth_sq = do_orth ? w.orth_sq : nullptr;
  q.E = w.E; q.P_max = t->p_max;
  q.tgt = tgt; q.n_desc = n_desc; q.child_w = t->child_w; q.proto_off = t->proto_off; q.cls_off = t->cls_off;
  q.V = V; q.V_first = V_first; q.N = t->n_nodes; q.P = t->n_protos; q.K = t->n_cols;
  q.eps = eps; q.mult = (flags & HCOMP_LOSS_SPARSITY) ? multiplier : 0.f; q.do_tanh = do_tanh; q.do_cls = do_cls;
  for (int i = 0; i < 4; ++i) q.lw.w[i] = weights_host[i];
  q.tanh_part = w.tanh_part; q.colsum = w.colsum; q.cls = w.cls; q.lse = ws + hcomp_head_losses_ws_floats(t);
  q.n_correct = n_correct; q.stats = stats; q.total = total; q.counter = counter;
  hc::head_chain_fwd_kernel<<<dim3(t->n_nodes, do_orth ? 4 : 3), 256, 0, S(stream)>>>(q);
  HC_LAUNCH_CHECK("head_chain_fwd");
  return 0;
}

int hcomp_head_chain_bwd(const float* g_total, const float* pooled, const float* out, const float* w_flat, const float* wc,
                         const int8_t* tgt, const int32_t* n_desc, const float* stats, const hcomp_tables* t, int V,
                         int V_first, int C, int flags, const float* weights_host, float eps, float multiplier,
                         const float* ws, const uint8_t* rel, float* g_pooled, float* g_wc, float* g_bias, float* g_align,
                         float* g_w, const int32_t* argmax, float thresh, const uint8_t* desc, int HW, void* scat_out,
                         float* coef_out, const hcomp_dz_blocks* blk, void* stream) {
  const LossWs w = loss_ws(const_cast<float*>(ws), t);
  const int N = t->n_nodes;
  if ((flags & HCOMP_LOSS_SPARSITY) && !(multiplier > 0.f)) return fail(HCOMP_E_ARG, "class loss: log1p(out**m) needs m > 0 (got %g)", multiplier);
  // the orth term's g_w (weights only) runs beside the chain kernel: fork BEFORE the chain launch (so the branch does not
  // depend on it), launch the chain kernel FIRST (the branch's P x C/256 blocks would otherwise queue in front of it)
  SideBranch* sb = nullptr;
  const bool orth_branch = g_w != nullptr && (flags & HCOMP_LOSS_ORTH);
  if (orth_branch) {
    if (int e = side_branch(&sb)) return e;
    HC_FORK(sb, S(stream));
  } else if (g_w != nullptr) {
    HC_CUDA(cudaMemsetAsync(g_w, 0, sizeof(float) * (size_t)t->n_protos * C, S(stream)));
  }
  hc::ChainBwdParams q{};
  q.g_total = g_total; q.pooled = pooled; q.out = out; q.wc = wc; q.colsum = w.colsum;
  q.lse = ws + hcomp_head_losses_ws_floats(t); q.tgt = tgt; q.n_desc = n_desc; q.child_w = t->child_w;
  q.proto_node = t->proto_node; q.proto_off = t->proto_off; q.cls_off = t->cls_off; q.wc_off = t->wc_off;
  q.col_node = t->col_node; q.welem_col = t->welem_col; q.welem_proto = t->welem_proto;
  q.V = V; q.V_first = V_first; q.N = N; q.P = t->n_protos; q.K = t->n_cols; q.n_w = t->n_welems;
  q.eps = eps; q.mult = (flags & HCOMP_LOSS_SPARSITY) ? multiplier : 0.f;
  q.do_tanh = (flags & HCOMP_LOSS_TANH) ? 1 : 0; q.do_cls = (flags & HCOMP_LOSS_CLASS) ? 1 : 0;
  for (int i = 0; i < 4; ++i) q.lw.w[i] = weights_host[i];
  q.g_pooled = g_pooled; q.g_wc = g_wc; q.g_bias = g_bias; q.g_align = g_align;
  q.nb_pooled = g_pooled ? blocks((long long)V * t->n_protos, 256) : 0;
  q.nb_wc = g_wc ? blocks(t->n_welems, 8) : 0;
  q.nb_bias = g_bias ? blocks(t->n_cols, 256) : 0;
  q.nb_align = g_align ? blocks(N, 256) : 0;
  int nb_coef = 0;
  if (scat_out != nullptr) {
    if (g_pooled == nullptr || argmax == nullptr) return fail(HCOMP_E_ARG, "chain_bwd: scat_out needs g_pooled and argmax");
    q.scat = reinterpret_cast<int2*>(scat_out); q.argmax = argmax; q.thresh = thresh;
  }
  if (coef_out != nullptr) {
    if (desc == nullptr) return fail(HCOMP_E_ARG, "chain_bwd: coef_out needs desc");
    q.coef = coef_out; q.desc = desc; q.HW = HW;
    nb_coef = blocks((long long)V_first * N, 256);
  }
  if (blk != nullptr && blk->t1 != nullptr && scat_out != nullptr) {      // marks go with the tables they describe
    if (blk->pcol == nullptr) return fail(HCOMP_E_ARG, "dZ block tables need pcol");
    if (g_align != nullptr && coef_out == nullptr) return fail(HCOMP_E_ARG, "dZ block tables: align gradient without coef_out");
    q.blk = dz_tables(blk, HW);
    q.HW = HW;
  }
  const int grid = q.nb_pooled + q.nb_wc + q.nb_bias + q.nb_align + nb_coef;
  if (grid > 0) {
    hc::head_chain_bwd_kernel<<<grid, 256, 0, S(stream)>>>(q);
    HC_LAUNCH_CHECK("head_chain_bwd");
  }
  if (orth_branch) {
    // (a shared-memory tiled variant -- one block per node x channel chunk, E_n and the kernel columns staged once -- was
    // measured at 10 us against this kernel's 8 us and removed; profiles/r2_k1_analysis.md section 7)
    hc::orth_bwd_scaled_kernel<<<dim3(t->n_protos, (C + 255) / 256), 256, 0, sb->stream>>>(
        w_flat, t->proto_node, t->proto_off, C, t->p_max, stats + 2 * N, w.E, rel, g_total, weights_host[2], g_w);
    HC_LAUNCH_CHECK("orth_bwd");
  }
  if (sb) HC_JOIN(sb, S(stream));
  return 0;
}

// ---------------------------------------------------------------- descendant-structured loss terms
namespace {
size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

hc::DescWs desc_ws(void* ws, const hcomp_tables* t, int V, size_t* total = nullptr) {
  hc::DescWs w;
  uint8_t* b = static_cast<uint8_t*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* p = b + off; off += align16(bytes); return p; };
  const size_t N = t->n_nodes, P = t->n_protos, K = t->n_cols, E = t->n_welems, VP = (size_t)V * P;
  w.acc = reinterpret_cast<double*>(take(sizeof(double) * 5 * N));
  w.leader = reinterpret_cast<int32_t*>(take(4 * (size_t)V));
  w.next = reinterpret_cast<int32_t*>(take(4 * (size_t)V));
  w.leaf_s1 = reinterpret_cast<float*>(take(4 * VP));
  w.leaf_s2 = reinterpret_cast<float*>(take(4 * VP));
  w.leaf_max = reinterpret_cast<float*>(take(4 * VP));
  w.leaf_arg = reinterpret_cast<int32_t*>(take(4 * VP));
  w.col_rel = reinterpret_cast<int32_t*>(take(4 * K));
  w.col_present = reinterpret_cast<int32_t*>(take(4 * K));
  w.cs_arg = reinterpret_cast<int32_t*>(take(4 * E));
  w.score = reinterpret_cast<float*>(take(4 * E));
  w.nz_prod = reinterpret_cast<float*>(take(4 * E));
  w.zeros = reinterpret_cast<int32_t*>(take(4 * E));
  w.y1_at = reinterpret_cast<float*>(take(4 * E));
  if (total) *total = off;
  return w;
}

hc::DescParams desc_params(const float* pooled, const float* wc, const float* presence, const float* gumbel,
                           const long long* ys, const int8_t* tgt, const int32_t* n_desc, const hcomp_tables* t, int V,
                           int V_first, int flags, const float* weights_host, float eps, float boost, float gumbel_tau) {
  hc::DescParams q;
  q.pooled = pooled; q.wc = wc; q.presence = presence; q.gumbel = gumbel; q.ys = ys; q.tgt = tgt; q.n_desc = n_desc;
  q.proto_off = t->proto_off; q.cls_off = t->cls_off; q.wc_off = t->wc_off; q.proto_node = t->proto_node;
  q.col_node = t->col_node; q.welem_col = t->welem_col; q.welem_proto = t->welem_proto; q.col_nleaves = t->col_nleaves;
  q.V = V; q.V_first = V_first; q.N = t->n_nodes; q.P = t->n_protos; q.K = t->n_cols; q.E = t->n_welems;
  q.flags = flags;
  q.w_td = weights_host[0]; q.w_cs = weights_host[1]; q.w_ov = weights_host[2]; q.w_l1 = weights_host[3];
  q.eps = eps; q.boost = boost; q.inv_tau = 1.0f / gumbel_tau;
  return q;
}

int desc_check(const float* presence, const float* gumbel, const hcomp_tables* t, int V, int flags, float gumbel_tau) {
  if (V <= 0 || V > 4096) return fail(HCOMP_E_ARG, "desc losses: batch of %d rows (supported: 1..4096)", V);
  if (t->col_nleaves == nullptr) return fail(HCOMP_E_ARG, "desc losses: hcomp_tables.col_nleaves missing");
  if ((flags & HCOMP_DESC_MASK_PRUNE) && (presence == nullptr || gumbel == nullptr || !(gumbel_tau > 0.f)))
    return fail(HCOMP_E_ARG, "mask pruning needs the presence logits, the Gumbel noise and tau > 0");
  return 0;
}
}  // namespace

long long hcomp_desc_losses_ws_bytes(const hcomp_tables* t, int V) {
  size_t total = 0;
  desc_ws(nullptr, t, V, &total);
  return (long long)total;
}

int hcomp_desc_losses_fwd(const float* pooled, const float* wc, const float* presence, const float* gumbel,
                          const long long* ys, const int8_t* tgt, const int32_t* n_desc, const hcomp_tables* t, int V,
                          int V_first, int flags, const float* weights_host, float eps, float boost, float gumbel_tau,
                          void* ws, float* stats, float* loss, void* stream) {
  if (int rc = desc_check(presence, gumbel, t, V, flags, gumbel_tau)) return rc;
  const hc::DescWs w = desc_ws(ws, t, V);
  const hc::DescParams q = desc_params(pooled, wc, presence, gumbel, ys, tgt, n_desc, t, V, V_first, flags, weights_host,
                                       eps, boost, gumbel_tau > 0.f ? gumbel_tau : 1.f);
  const bool td = flags & HCOMP_DESC_TANH_DESC, cs = flags & HCOMP_DESC_CONTRAST, mp = flags & HCOMP_DESC_MASK_PRUNE;
  HC_CUDA(cudaMemsetAsync(w.acc, 0, sizeof(double) * 5 * (size_t)t->n_nodes, S(stream)));
  hc::desc_prep_kernel<<<1, 256, sizeof(long long) * V, S(stream)>>>(ys, V, w.leader, w.next);
  HC_LAUNCH_CHECK("desc_prep");
  if (td || mp) {
    hc::desc_leaf_stats_kernel<<<dim3((t->n_protos + 127) / 128, V), 128, 0, S(stream)>>>(q, w);
    HC_LAUNCH_CHECK("desc_leaf_stats");
  }
  hc::desc_col_stats_kernel<<<(t->n_cols * 32 + 127) / 128, 128, 0, S(stream)>>>(q, w);
  HC_LAUNCH_CHECK("desc_col_stats");
  if (cs || mp) {
    hc::desc_elem_reduce_kernel<<<(int)(((long long)t->n_welems * 32 + 127) / 128), 128, 0, S(stream)>>>(q, w);
    HC_LAUNCH_CHECK("desc_elem_reduce");
  }
  if (td) {
    hc::tanh_desc_fwd_kernel<<<dim3((t->n_nodes + 63) / 64, V), 64, 0, S(stream)>>>(q, w);
    HC_LAUNCH_CHECK("tanh_desc_fwd");
  }
  if (mp) {
    hc::mask_prune_fwd_kernel<<<(t->n_protos + 127) / 128, 128, 0, S(stream)>>>(q, w);
    HC_LAUNCH_CHECK("mask_prune_fwd");
  }
  hc::desc_combine_kernel<<<1, 256, 0, S(stream)>>>(q, w, stats, loss);
  HC_LAUNCH_CHECK("desc_combine");
  return 0;
}

int hcomp_desc_losses_bwd(const float* g_loss, const float* pooled, const float* wc, const float* presence,
                          const float* gumbel, const long long* ys, const int8_t* tgt, const int32_t* n_desc,
                          const hcomp_tables* t, int V, int V_first, int flags, const float* weights_host, float eps,
                          float boost, float gumbel_tau, const void* ws, float* g_pooled, float* g_presence, void* stream) {
  if (int rc = desc_check(presence, gumbel, t, V, flags, gumbel_tau)) return rc;
  const hc::DescWs w = desc_ws(const_cast<void*>(ws), t, V);
  const hc::DescParams q = desc_params(pooled, wc, presence, gumbel, ys, tgt, n_desc, t, V, V_first, flags, weights_host,
                                       eps, boost, gumbel_tau > 0.f ? gumbel_tau : 1.f);
  if (g_pooled != nullptr) {
    hc::desc_bwd_pooled_kernel<<<dim3((t->n_protos + 127) / 128, V), 128, 0, S(stream)>>>(q, w, g_loss, g_pooled);
    HC_LAUNCH_CHECK("desc_bwd_pooled");
  }
  if (g_presence != nullptr) {
    if (flags & HCOMP_DESC_MASK_PRUNE) {
      hc::mask_prune_bwd_presence_kernel<<<(t->n_protos + 127) / 128, 128, 0, S(stream)>>>(q, w, g_loss, g_presence);
      HC_LAUNCH_CHECK("mask_prune_bwd_presence");
    } else {
      HC_CUDA(cudaMemsetAsync(g_presence, 0, sizeof(float) * 2 * (size_t)t->n_protos, S(stream)));
    }
  }
  return 0;
}

int hcomp_joint_leaf(const float* out, const hcomp_tables* t, int V, float tau, const float* prob_override,
                     float* probs_ws, float* joint, long long* pred, void* stream) {
  if (!(tau > 0.f)) return fail(HCOMP_E_ARG, "path-probability tau must be > 0");
  hc::node_probs_kernel<<<blocks((long long)V * t->n_nodes, 128), 128, 0, S(stream)>>>(out, t->cls_off, V, t->n_nodes,
                                                                                      t->n_cols, 1.f / tau, prob_override,
                                                                                      probs_ws);
  HC_LAUNCH_CHECK("node_probs");
  hc::leaf_joint_kernel<<<blocks((long long)V * t->n_leaves, 128), 128, 0, S(stream)>>>(probs_ws, t->path_off, t->path_col,
                                                                                       V, t->n_leaves, t->n_cols, joint);
  HC_LAUNCH_CHECK("leaf_joint");
  if (pred) {
    hc::row_argmax_kernel<<<blocks(V, 128), 128, 0, S(stream)>>>(joint, V, t->n_leaves, pred);
    HC_LAUNCH_CHECK("row_argmax");
  }
  return 0;
}

int hcomp_topk_update(const float* pooled, const int32_t* argmax, const long long* ys, const long long* img_ids,
                      const float* wc, const hcomp_tables* t, int V, int k, int find_non_descendants, int32_t* ws_2v,
                      float* t_score, long long* t_img, int32_t* t_loc, void* stream) {
  if (k <= 0 || k > hc::TOPK_MAX) return fail(HCOMP_E_ARG, "topk=%d (supported: 1..%d)", k, hc::TOPK_MAX);
  if (V <= 0 || V > 4096) return fail(HCOMP_E_ARG, "topk update: batch of %d rows (supported: 1..4096)", V);
  int32_t* leader = ws_2v;
  int32_t* next = ws_2v + V;
  hc::desc_prep_kernel<<<1, 256, sizeof(long long) * V, S(stream)>>>(ys, V, leader, next);
  HC_LAUNCH_CHECK("desc_prep");
  hc::topk_update_kernel<<<dim3((t->n_protos + 127) / 128, V), 128, 0, S(stream)>>>(
      pooled, argmax, ys, img_ids, leader, next, t->anc, t->proto_node, t->proto_off, t->cls_off, t->wc_off, wc,
      find_non_descendants, V, t->n_protos, t->n_nodes, t->n_leaves, k, t_score, t_img, t_loc);
  HC_LAUNCH_CHECK("topk_update");
  return 0;
}

int hcomp_materialize_map(const void* x_bf16, const float* w_node, int V, int HW, int C, int P_n, float tau, float* map,
                          void* stream) {
  const int warps = 8;
  const long long rows = (long long)V * HW;
  hc::materialize_map_kernel<<<blocks(rows, warps), warps * 32, warps * P_n * sizeof(float), S(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x_bf16), w_node, V, HW, C, P_n, 1.f / tau, map);
  HC_LAUNCH_CHECK("materialize_map");
  return 0;
}

#ifdef HC_EXP_TIMING
/* timing experiment only: read and clear the per-role cycle counters of the fused pair kernels */
int hcomp_debug_pair_counters(unsigned long long* out16) {
  HC_CUDA(cudaDeviceSynchronize());
  HC_CUDA(cudaMemcpyFromSymbol(out16, hc::g_pair_dbg, sizeof(unsigned long long) * 16));
  unsigned long long z[16] = {0};
  HC_CUDA(cudaMemcpyToSymbol(hc::g_pair_dbg, z, sizeof(z)));
  return 0;
}
/* per-CTA wall-clock stamps of the LAST pair-kernel launch: out[160*4] ns */
int hcomp_debug_pair_stamps(unsigned long long* out640) {
  HC_CUDA(cudaDeviceSynchronize());
  HC_CUDA(cudaMemcpyFromSymbol(out640, hc::g_pair_stamps, sizeof(unsigned long long) * 640));
  return 0;
}
/* per-item event trace of four CTAs of the LAST pair-kernel launch: out[4*16*8] ns (cleared after the read) */
int hcomp_debug_pair_trace(unsigned long long* out512) {
  HC_CUDA(cudaDeviceSynchronize());
  HC_CUDA(cudaMemcpyFromSymbol(out512, hc::g_pair_trace, sizeof(unsigned long long) * 512));
  static unsigned long long z[512];
  HC_CUDA(cudaMemcpyToSymbol(hc::g_pair_trace, z, sizeof(z)));
  return 0;
}
/* all epilogue warps of CTA 0 of the LAST pair-kernel launch: out[12*16*8] ns (cleared after the read) */
int hcomp_debug_pair_wtrace(unsigned long long* out1536) {
  HC_CUDA(cudaDeviceSynchronize());
  HC_CUDA(cudaMemcpyFromSymbol(out1536, hc::g_pair_wtrace, sizeof(unsigned long long) * 1536));
  static unsigned long long z[1536];
  HC_CUDA(cudaMemcpyToSymbol(hc::g_pair_wtrace, z, sizeof(z)));
  return 0;
}
#endif

int hcomp_allreduce_mean_symm(float* local, float* mc, const void* peers_dev, const void* pads_dev, int rank, int world,
                              long long n, int channel_base, int ctas, void* stream) {
  if (world < 2 || rank < 0 || rank >= world) return fail(HCOMP_E_ARG, "all-reduce: rank %d of %d", rank, world);
  if (n <= 0 || n % 4 != 0 || (reinterpret_cast<uintptr_t>(local) & 15) != 0)
    return fail(HCOMP_E_ARG, "all-reduce: n=%lld floats must be a positive multiple of 4 and the buffer 16-byte aligned", n);
  if (pads_dev == nullptr || (mc == nullptr && peers_dev == nullptr)) return fail(HCOMP_E_ARG, "all-reduce: missing signal pads / peer pointers");
  if (ctas < 1 || ctas > 64 || world > 256) return fail(HCOMP_E_ARG, "all-reduce: ctas=%d world=%d", ctas, world);
  hc::AllreduceParams p;
  p.local = local; p.mc = mc;
  p.peers = reinterpret_cast<float* const*>(peers_dev);
  p.pads = reinterpret_cast<uint32_t* const*>(pads_dev);
  p.rank = rank; p.world = world; p.n = n; p.channel_base = channel_base; p.scale = 1.f / float(world);
  hc::allreduce_mean_kernel<<<ctas, 256, 0, S(stream)>>>(p);
  HC_LAUNCH_CHECK("allreduce_mean");
  return 0;
}

int hcomp_gemm_bf16(const void* a, const void* b, int M, int N, int K, int a_mn, int b_mn, int out_mode, int splits,
                    void* out, long long ldo, void* stream) {
  return run_gemm(a, b, M, N, K, a_mn != 0, b_mn != 0, out_mode, splits, out, ldo, nullptr, S(stream));
}

}  // extern "C"
