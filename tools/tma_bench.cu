// Microbenchmark: per-SM TMA ingest rate.  Every CTA (1 per SM) streams [128 rows x 64 bf16] boxes (16 KB, 128B swizzle)
// of a row-major [rows, 768] bf16 matrix into a 4-stage x 48 KB shared-memory ring exactly like the K1 producer does
// (3 boxes per stage), a consumer thread releases each stage as soon as it has landed.  Reports bytes/clk/SM and TB/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pipnet_b200/csrc -o tools/tma_bench tools/tma_bench.cu -lcuda
#include <cstdio>
#include <cuda.h>
#include "ptx.cuh"
using namespace hc;

__global__ void __launch_bounds__(128, 1) tma_stream(const __grid_constant__ CUtensorMap map, int rows_total, int iters,
                                                     int boxes_per_stage, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[4], empty[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } fence_mbar_init(); }
  __syncthreads();
  const int tiles = rows_total / 128;
  long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    int tile = blockIdx.x;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_arrive_expect_tx(&full[stage], boxes_per_stage * 16384);
      for (int b = 0; b < boxes_per_stage; ++b) {
        tma_load_2d(smem + stage * 49152 + b * 16384, &map, &full[stage], (it % 12) * 64, (tile % tiles) * 128);
        tile += (b == boxes_per_stage - 1 && (it % 12) == 11) ? gridDim.x : 0;
        if (b < boxes_per_stage - 1) tile += 0;
      }
      if (++stage == 4) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[stage], phase);
      mbar_arrive(&empty[stage]);
      if (++stage == 4) { stage = 0; phase ^= 1; }
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
  cudaFuncSetAttribute(tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (long long rows : {8192LL, 86528LL, 173056LL}) {          // 12.6 MB (L2 resident), 133 MB, 266 MB
    void* buf; cudaMalloc(&buf, rows * 768 * 2); cudaMemset(buf, 0, rows * 768 * 2);
    CUtensorMap m; cuuint64_t gd[2] = {768, (cuuint64_t)rows}; cuuint64_t gs[1] = {1536}; cuuint32_t bx[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
    enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    for (int grid : {18, 37, 74, 148}) { const int boxes = 3;
      const int iters = 12 * 40;
      tma_stream<<<grid, 128, 200 * 1024>>>(m, (int)rows, 24, boxes, d); cudaDeviceSynchronize();
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      tma_stream<<<grid, 128, 200 * 1024>>>(m, (int)rows, iters, boxes, d);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      const double bytes = double(iters) * boxes * 16384;
      printf("array %6.1f MB  CTAs %3d : %6.1f B/clk/SM  %6.2f TB/s chip  (%.1f us) %s\n", rows * 1536 / 1e6, grid,
             bytes / double(cyc), bytes * grid / (ms * 1e-3) / 1e12, ms * 1e3, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    cudaFree(buf);
  }
  return 0;
}
