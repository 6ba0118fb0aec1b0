// Microbenchmark 3: where do the ~1000 cycles per ring iteration of tools/tma_bench2.cu go?  Same producer / consumer
// ring, but every operation of both loops is bracketed with clock64() and the per-operation averages are printed:
//   producer: wait(empty) | arrive.expect_tx | TMA issue        consumer: wait(full) | arrive(empty)
// Variants: try_wait (HW suspend) vs test_wait (pure polling) on either side, tensor-map prefetch, consumer = tcgen05
// commit is NOT modelled here (plain arrive).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pipnet_b200/csrc -o tools/tma_bench3 tools/tma_bench3.cu -lcuda
#include <cstdio>
#include <cuda.h>
#include "ptx.cuh"
using namespace hc;

constexpr int MAX_STAGES = 32;

template <bool POLL>
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  if (POLL) { while (!mbar_test_wait(bar, parity)) {} }
  else { while (!mbar_try_wait(bar, parity)) {} }
}

template <bool POLL_P, bool POLL_C>
__global__ void __launch_bounds__(128, 1) tma_stream(const __grid_constant__ CUtensorMap map, int rows_total, int iters,
                                                     int n_stages, int boxes_per_stage, int box_rows, int prefetch,
                                                     long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[MAX_STAGES], empty[MAX_STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int box_bytes = box_rows * 128;
  const int stage_bytes = boxes_per_stage * box_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_mbar_init();
    if (prefetch) prefetch_tmap(&map);
  }
  __syncthreads();
  const int tiles = rows_total / box_rows;
  long long t0 = clock64();
  long long a0 = 0, a1 = 0, a2 = 0;
  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    int tile = blockIdx.x, kc = 0;
    for (int it = 0; it < iters; ++it) {
      long long c0 = clock64();
      wait_bar<POLL_P>(&empty[stage], phase ^ 1);
      long long c1 = clock64();
      mbar_arrive_expect_tx(&full[stage], stage_bytes);
      long long c2 = clock64();
      for (int b = 0; b < boxes_per_stage; ++b) {
        int t = tile + b * 7; if (t >= tiles) t -= tiles;
        tma_load_2d(smem + stage * stage_bytes + b * box_bytes, &map, &full[stage], kc * 64, t * box_rows);
      }
      long long c3 = clock64();
      a0 += c1 - c0; a1 += c2 - c1; a2 += c3 - c2;
      if (++kc == 12) { kc = 0; tile += gridDim.x; if (tile >= tiles) tile -= tiles; }
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
    if (blockIdx.x == 0) { out[1] = a0; out[2] = a1; out[3] = a2; }
  } else if (warp == 3 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      long long c0 = clock64();
      wait_bar<POLL_C>(&full[stage], phase);
      long long c1 = clock64();
      mbar_arrive(&empty[stage]);
      long long c2 = clock64();
      a0 += c1 - c0; a1 += c2 - c1;
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
    if (blockIdx.x == 0) { out[4] = a0; out[5] = a1; }
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}

typedef void (*KernT)(const CUtensorMap, int, int, int, int, int, int, long long*);

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  Enc enc = (Enc)fp;
  long long* d; cudaMalloc(&d, 64);
  const int SMEM = 225 * 1024;
  KernT kerns[4] = {tma_stream<false, false>, tma_stream<true, false>, tma_stream<false, true>, tma_stream<true, true>};
  const char* kn[4] = {"try/try", "pollP/tryC", "tryP/pollC", "poll/poll"};
  for (auto k : kerns) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  const long long rows = 86528;
  void* buf; cudaMalloc(&buf, rows * 768 * 2); cudaMemset(buf, 0, rows * 768 * 2);
  struct Cfg { int stages, boxes, box_rows, prefetch; };
  const Cfg cfgs[] = {{4, 3, 128, 0}, {12, 1, 128, 0}, {12, 1, 128, 1}, {5, 3, 128, 1}, {4, 1, 128, 1}};
  for (const Cfg& c : cfgs) {
    CUtensorMap m; cuuint64_t gd[2] = {768, (cuuint64_t)rows}; cuuint64_t gs[1] = {1536};
    cuuint32_t bx[2] = {64, (cuuint32_t)c.box_rows}; cuuint32_t es[2] = {1, 1};
    enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int stage_bytes = c.boxes * c.box_rows * 128;
    if (c.stages * stage_bytes + 2048 > SMEM) { printf("skip (smem)\n"); continue; }
    const int iters = (12 * 40 * 49152) / stage_bytes;
    for (int v = 0; v < 4; ++v) {
      for (int grid : {148}) {
        kerns[v]<<<grid, 128, SMEM>>>(m, (int)rows, 24, c.stages, c.boxes, c.box_rows, c.prefetch, d);
        cudaDeviceSynchronize();
        kerns[v]<<<grid, 128, SMEM>>>(m, (int)rows, iters, c.stages, c.boxes, c.box_rows, c.prefetch, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[6]; cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
        const double bytes = double(iters) * stage_bytes;
        printf("%-11s stages %2d x %d boxes (%3d KB) prefetch %d : %5.1f B/clk/SM, %6.0f cyc/iter | producer: wait %6.0f expect_tx %5.0f tma %5.0f | consumer: wait %6.0f arrive %5.0f  %s\n",
               kn[v], c.stages, c.boxes, c.stages * stage_bytes / 1024, c.prefetch, bytes / double(h[0]), double(h[0]) / iters,
               double(h[1]) / iters, double(h[2]) / iters, double(h[3]) / iters, double(h[4]) / iters, double(h[5]) / iters,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
