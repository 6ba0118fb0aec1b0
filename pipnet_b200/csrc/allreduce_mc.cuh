// Mean all-reduce of the head's flat fp32 gradient bucket over one NVSwitch box, written for the way this path uses it:
// ONE small latency-bound buffer per step (cub27: 1.5 MB, cub190: 11.6 MB) that has to be reduced while the dX GEMM
// owns the GPU.  Replaces, on the data-parallel path of SURVEY.md section 8e (main_dist.py:330 semantics: mean of the
// per-rank gradients), the NCCL ring kernel whose ~22-37 us tail was exposed behind the persistent GEMM (round-1
// SCALE): a handful of CTAs on SMs the dX GEMM leaves free (hcomp_set_reserved_sms) do
//     barrier  ->  reduce my 1/W shard IN THE SWITCH (multimem.ld_reduce)  ->  broadcast it (multimem.st)  ->  barrier
// over a symmetric-memory bucket (every rank's buffer mapped at one multicast address).  Without multicast support
// the same kernel runs the two-shot exchange with plain peer loads / stores.
//
// Cross-rank barrier = the signal-pad protocol of torch's symmetric memory (one 32-bit flag per (channel, source
// rank) in every rank's pad): put = CAS 0->1 with release.sys on the PEER's pad, wait = CAS 1->0 with acquire.sys on
// my own; flags return to 0, so the kernel is replayable inside a CUDA graph.  Every CTA runs its own barrier on its
// own channel -- no grid-wide synchronisation inside the kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hc {

struct AllreduceParams {
  float* local;               // this rank's bucket (symmetric allocation)
  float* mc;                  // multicast address of the bucket or nullptr
  float* const* peers;        // [world] device pointers to every rank's bucket (peer-mapped), used without multicast
  uint32_t* const* pads;      // [world] device pointers to every rank's signal pad
  int rank, world;
  long long n;                // floats; multiple of 4
  int channel_base;           // first signal-pad channel of this call (one channel per CTA)
  float scale;                // 1 / world
};

__device__ __forceinline__ void ar_put_signal(uint32_t* addr) {
  uint32_t old;
  do {
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
  } while (old != 0u);
}
__device__ __forceinline__ void ar_wait_signal(uint32_t* addr) {
  uint32_t old;
  do {
    asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory");
  } while (old != 1u);
}
// all ranks' CTA `channel` meet here; memory operations before it (any thread of the CTA) are visible system-wide after it
__device__ __forceinline__ void ar_barrier(const AllreduceParams& p, int channel) {
  __syncthreads();
  if (threadIdx.x < p.world) {
    __threadfence_system();
    const int peer = threadIdx.x;
    ar_put_signal(p.pads[peer] + (size_t)channel * p.world + p.rank);
    ar_wait_signal(p.pads[p.rank] + (size_t)channel * p.world + peer);
  }
  __syncthreads();
}

__device__ __forceinline__ float4 mm_ld_reduce_add(const float* mc_addr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc_addr) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(float* mc_addr, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// grid = a few CTAs (each owns a strided part of this rank's shard), 256 threads
__global__ void __launch_bounds__(256) allreduce_mean_kernel(const AllreduceParams p) {
  const int channel = p.channel_base + blockIdx.x;
  // shard of this rank in float4 units, split over the CTAs of the grid
  const long long n4 = p.n >> 2;
  const long long per = (n4 + p.world - 1) / p.world;
  const long long lo = per * p.rank, hi = (lo + per < n4) ? lo + per : n4;
  ar_barrier(p, channel);                       // every rank's producers (dW GEMM, ...) are done: stream order + this
  if (p.mc != nullptr) {
    // UNROLL in-switch reductions in flight per thread before the first store: one multimem.ld_reduce is a ~2-3 us round
    // trip through the switch, so a load -> store chain per element left the links idle (cub190's 11.7 MB bucket took
    // 264 us = 44 GB/s on 8 CTAs at 2 GPUs)
    constexpr int UNROLL = 8;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * UNROLL) {
      float4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (i0 + u * stride < hi) v[u] = mm_ld_reduce_add(p.mc + 4 * (i0 + u * stride));       // sum over ranks, computed by the switch
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (i0 + u * stride < hi) {
          v[u].x *= p.scale; v[u].y *= p.scale; v[u].z *= p.scale; v[u].w *= p.scale;
          mm_st(p.mc + 4 * (i0 + u * stride), v[u]);                                           // lands in every rank's bucket
        }
      }
    }
  } else {
    for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < p.world; ++r) {                 // fixed rank order: every rank computes bit-identical sums
        const float4 v = __ldcv(reinterpret_cast<const float4*>(p.peers[r]) + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      acc.x *= p.scale; acc.y *= p.scale; acc.z *= p.scale; acc.w *= p.scale;
      for (int r = 0; r < p.world; ++r) __stcg(reinterpret_cast<float4*>(p.peers[r]) + i, acc);
    }
  }
  ar_barrier(p, channel);                       // every shard has landed everywhere
}

}  // namespace hc
