// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, SS mode, K-major 128B-swizzled operands resident in
// shared memory) for different N, with one or two accumulators, issued back to back by one thread of one CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pipnet_b200/csrc -o tools/mma_bench tools/mma_bench.cu
#include <cstdio>
#include "ptx.cuh"
using namespace hc;

template <int N, int NACC, int DISTINCT_OPERANDS>
__global__ void __launch_bounds__(128, 1) mma_rate(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0 && lane == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc<512>(&tbase);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    constexpr uint32_t HI = desc_hi32(1024);
    constexpr uint32_t LOF = desc_lo_flags(16);
    const uint32_t idesc = make_idesc(128, N, false, false);
    const uint32_t base = (smem_u32(smem) >> 4) | LOF;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // A tile: 16 KB blocks, B tile: 32 KB blocks; optionally walk over distinct smem regions like a real k loop
          const uint32_t a = base + ((DISTINCT_OPERANDS ? ((it & 1) * 49152 + (k >> 2) * 0) : 0) >> 4) + 2 * (k & 3);
          const uint32_t b = base + ((16384 + (DISTINCT_OPERANDS ? (it & 1) * 49152 : 0)) >> 4) + 2 * (k & 3);
          umma_bf16(tbase + (NACC == 2 ? (k & 1) * 256 : 0), desc64(a, HI), desc64(b, HI), idesc, 1u);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (elect_one() && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tbase); }
}

template <int N, int NACC, int D>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 8);
  auto k = mma_rate<N, NACC, D>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  k<<<grid, 128, 200 * 1024>>>(d, 10); cudaDeviceSynchronize();
  k<<<grid, 128, 200 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-34s grid %3d  N=%3d acc=%d distinct=%d : %7.1f cycles/MMA  (ideal %d)  %s\n", name, grid, N, NACC, D,
         double(h) / (iters * 8.0), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  for (int grid : {1, 148}) {
    run<64, 1, 0>("128xNx16 same operands", grid);
    run<128, 1, 0>("128xNx16 same operands", grid);
    run<128, 2, 0>("128xNx16 same operands, 2 acc", grid);
    run<256, 1, 0>("128xNx16 same operands", grid);
    run<256, 2, 0>("128xNx16 same operands, 2 acc", grid);
    run<128, 2, 1>("128xNx16 alternating stages, 2 acc", grid);
    run<256, 2, 1>("128xNx16 alternating stages, 2 acc", grid);
  }
  return 0;
}
