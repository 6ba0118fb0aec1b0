// Microbenchmark 2: what limits per-SM TMA ingest?  Sweeps ring granularity (stages x boxes per stage), box height
// (64 / 128 / 256 rows of 128 bytes), the number of producer threads and the L2 promotion size at (nearly) constant
// bytes in flight.  One CTA per SM streams 128B-swizzled boxes of a row-major [rows, 768] bf16 matrix, a consumer
// thread releases each stage as soon as it has landed (as tools/tma_bench.cu does).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pipnet_b200/csrc -o tools/tma_bench2 tools/tma_bench2.cu -lcuda
#include <cstdio>
#include <cuda.h>
#include "ptx.cuh"
using namespace hc;

constexpr int MAX_STAGES = 32;

__global__ void __launch_bounds__(128, 1) tma_stream(const __grid_constant__ CUtensorMap map, int rows_total, int iters,
                                                     int n_stages, int boxes_per_stage, int box_rows, int producers,
                                                     long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[MAX_STAGES], empty[MAX_STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int box_bytes = box_rows * 128;
  const int stage_bytes = boxes_per_stage * box_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) { mbar_init(&full[i], producers); mbar_init(&empty[i], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int tiles = rows_total / box_rows;
  long long t0 = clock64();
  if (warp < producers && lane == 0) {
    // producer w issues boxes w, w+producers, ... of every stage
    int stage = 0; uint32_t phase = 0;
    int tile = blockIdx.x;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&empty[stage], phase ^ 1);
      int mine = 0;
      for (int b = warp; b < boxes_per_stage; b += producers) ++mine;
      mbar_arrive_expect_tx(&full[stage], mine * box_bytes);
      for (int b = warp; b < boxes_per_stage; b += producers)
        tma_load_2d(smem + stage * stage_bytes + b * box_bytes, &map, &full[stage], (it % 12) * 64,
                    ((tile + b * 7) % tiles) * box_rows);
      if ((it % 12) == 11) tile += gridDim.x;
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 3 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[stage], phase);
      mbar_arrive(&empty[stage]);
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  Enc enc = (Enc)fp;
  long long* d; cudaMalloc(&d, 8);
  const int SMEM = 225 * 1024;
  cudaFuncSetAttribute(tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  const long long rows = 86528;      // 133 MB
  void* buf; cudaMalloc(&buf, rows * 768 * 2); cudaMemset(buf, 0, rows * 768 * 2);
  struct Cfg { int stages, boxes, box_rows, producers, promo; };
  const Cfg cfgs[] = {
      {4, 3, 128, 1, 256},    // round-1 baseline: 192 KB in flight, 48 KB stages
      {4, 1, 128, 1, 256},    // 64 KB
      {8, 1, 128, 1, 256},    // 128 KB, 16 KB stages
      {12, 1, 128, 1, 256},   // 192 KB, 16 KB stages
      {13, 1, 128, 1, 256},   // 208 KB
      {6, 2, 128, 1, 256},    // 192 KB, 32 KB stages
      {5, 5, 64, 1, 256},     // 200 KB, 40 KB stages of 8 KB boxes
      {24, 1, 64, 1, 256},    // 192 KB, 8 KB stages
      {6, 1, 256, 1, 256},    // 192 KB, 32 KB boxes
      {4, 3, 128, 3, 256},    // three producer threads
      {12, 1, 128, 1, 128},   // L2 promotion 128 B
      {12, 1, 128, 1, 0},     // no L2 promotion
      {5, 5, 64, 1, 128},
  };
  for (const Cfg& c : cfgs) {
    CUtensorMap m; cuuint64_t gd[2] = {768, (cuuint64_t)rows}; cuuint64_t gs[1] = {1536};
    cuuint32_t bx[2] = {64, (cuuint32_t)c.box_rows}; cuuint32_t es[2] = {1, 1};
    CUtensorMapL2promotion pr = c.promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                              : c.promo == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int stage_bytes = c.boxes * c.box_rows * 128;
    if (c.stages * stage_bytes + 2048 > SMEM) { printf("skip (smem)\n"); continue; }
    const int iters = (12 * 40 * 49152) / stage_bytes;     // same total bytes for every configuration
    for (int grid : {37, 148}) {
      tma_stream<<<grid, 128, SMEM>>>(m, (int)rows, 24, c.stages, c.boxes, c.box_rows, c.producers, d);
      cudaDeviceSynchronize();
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      tma_stream<<<grid, 128, SMEM>>>(m, (int)rows, iters, c.stages, c.boxes, c.box_rows, c.producers, d);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      const double bytes = double(iters) * stage_bytes;
      printf("stages %2d x %d boxes of %3d rows (%3d KB in flight, %d producer(s), promo %3d)  CTAs %3d : %6.1f B/clk/SM  %6.2f TB/s chip  %s\n",
             c.stages, c.boxes, c.box_rows, c.stages * stage_bytes / 1024, c.producers, c.promo, grid, bytes / double(cyc),
             bytes * grid / (ms * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
