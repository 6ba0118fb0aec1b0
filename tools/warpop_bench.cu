// Microbenchmark: issue rate of the warp-level instructions the K1 pooling epilogue is made of, per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/warpop_bench tools/warpop_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(uint32_t* out, long long* cyc, int iters) {
  uint32_t v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x;
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = float(v[i] & 1023) * 1e-3f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) { uint32_t r; asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v[i])); v[i] = r + i; }
      if (OP == 1) { uint32_t r; asm volatile("{ .reg .pred p; setp.ne.u32 p, %1, 0; vote.sync.ballot.b32 %0, p, 0xffffffff; }" : "=r"(r) : "r"(v[i])); v[i] = r ^ (v[i] + 1); }
      if (OP == 2) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i])); f[i] = f[i] * 0.5f; }
      if (OP == 3) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(0.999f)); }
      if (OP == 4) { uint32_t r; asm volatile("shfl.sync.bfly.b32 %0, %1, 1, 0x1f, 0xffffffff;" : "=r"(r) : "r"(v[i])); v[i] = r + 1; }
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += v[i] + __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, uint32_t* out, long long* cyc) {
  for (int warps : {4, 12, 16}) {
    const int iters = 2000;
    k<OP><<<148, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_smsp = double(iters) * 8 * (warps / 4.0);      // warp-instructions per sub-partition
    printf("%-22s %2d warps/SM: %6.2f cycles per warp-instruction per SMSP (incl. the dependent ALU op)\n", name, warps, double(c) / per_smsp);
  }
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  run<0>("redux.sync.max.u32", out, cyc);
  run<1>("setp+vote.ballot", out, cyc);
  run<2>("ex2.approx (+fmul)", out, cyc);
  run<3>("fma.rn.f32", out, cyc);
  run<4>("shfl.bfly (+add)", out, cyc);
  return 0;
}
