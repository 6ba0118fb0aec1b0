// Warp-specialised tcgen05 GEMM, D[M,N] (+)= A[M,K] * B[K,N], bf16 in / fp32 accumulate in TMEM.
// Used for the head's two backward contractions (reference: autograd of the per-node 1x1 convs,
// pipnet/train.py:264 -> cuDNN dgrad/wgrad per node):
//   dX[rows, C]  = dZ[rows, P_c] * Wpc[P_c, C]        A K-major , B MN-major      (P_c: compact dZ column axis)
//   dW[P_c, C]  += dZ^T[P_c, rows] * X[rows, C]       A MN-major, B MN-major, split-K + fp32 red.add (rows -> flat via row_map_c)
// Tile 128 x 256 x 64, 4-stage TMA ring (48 KB / stage), 2 accumulator stages of 256 TMEM columns.
// "MN-major" operands are consumed straight from row-major storage whose contiguous axis is the
// M (or N) dimension -- no transposed copies in HBM.
#pragma once
#include "ptx.cuh"

namespace hc {

constexpr int G_BM = 128, G_BN = 256, G_BK = 64;
constexpr int G_A_BYTES = G_BM * G_BK * 2;   // 16 KB
constexpr int G_B_BYTES = G_BN * G_BK * 2;   // 32 KB
constexpr int G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;
constexpr int G_THREADS = 384;
constexpr int G_OUT_STAGE_BYTES = G_BM * G_BN * 2;       // bf16 output tile staged for TMA stores (64 KB)
// bf16 output: 3 operand stages + the 64 KB store staging; otherwise 4 operand stages
template <int OUT> struct GemmCfg {
  static constexpr int STAGES = (OUT == 0) ? 3 : 4;
  static constexpr int SMEM_BYTES = STAGES * G_STAGE_BYTES + (OUT == 0 ? G_OUT_STAGE_BYTES : 0) + 1024 + 256;
};
constexpr int G_MAX_STAGES = 4;

enum GemmOut : int { OUT_BF16 = 0, OUT_F32 = 1, OUT_RED_F32 = 2 };

struct GemmParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles, num_k_blocks, splits, k_blocks_per_split;
  void* out;
  long long ldo;
  const int32_t* row_map;   // OUT_RED_F32: GEMM row -> output row (or -1 to drop); may be null (identity)
  // optional (CTA-pair kernels): kact[mt2 * kact_ld + kb] == 0 says that the A operand's 256-row x 64-k block (mt2, kb) is
  // all zeros, so the k-block is neither fetched nor multiplied (block-sparse dZ, see hcomp_head_bwd_dx).  Rows of the
  // table are 8-byte aligned (kact_ld % 8 == 0) and padded with zeros.
  const uint8_t* kact;
  int kact_ld;
};

struct GemmSmem {
  uint64_t full[G_MAX_STAGES];
  uint64_t empty[G_MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool A_MN, bool B_MN, int OUT>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_o, const GemmParams p) {
  constexpr int G_STAGES = GemmCfg<OUT>::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ostage = smem + G_STAGES * G_STAGE_BYTES;        // OUT_BF16 only: 4 boxes [128 rows x 64 cols], 128B-swizzled
  GemmSmem* sb = reinterpret_cast<GemmSmem*>(ostage + (OUT == OUT_BF16 ? G_OUT_STAGE_BYTES : 0));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_mn = p.num_m_tiles * p.num_n_tiles;
  const int total_items = tiles_mn * p.splits;
  const int worker = blockIdx.x, num_workers = gridDim.x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < G_STAGES; ++i) {
      mbar_init(&sb->full[i], 1);
      mbar_init(&sb->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sb->tmem_full[i], 1);
      mbar_init(&sb->tmem_empty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(&sb->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sb->tmem_base;

  // item -> (split, m tile, n tile); n fastest so neighbouring CTAs share the A tile in L2
  auto decode = [&](int item, int& mt, int& nt, int& kb0, int& kb1) {
    const int sp = item / tiles_mn;
    const int r = item - sp * tiles_mn;
    mt = r / p.num_n_tiles;
    nt = r - mt * p.num_n_tiles;
    kb0 = sp * p.k_blocks_per_split;
    kb1 = min(p.num_k_blocks, kb0 + p.k_blocks_per_split);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = worker; item < total_items; item += num_workers) {
        int mt, nt, kb0, kb1;
        decode(item, mt, nt, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&sb->empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * G_STAGE_BYTES;
          uint8_t* sbm = sa + G_A_BYTES;
          mbar_arrive_expect_tx(&sb->full[stage], G_STAGE_BYTES);
          if constexpr (!A_MN) {
            tma_load_2d(sa, &tmap_a, &sb->full[stage], kb * G_BK, mt * G_BM);            // box [64 k, 128 m]
          } else {
#pragma unroll
            for (int c = 0; c < G_BM / 64; ++c)                                            // box [64 m, 64 k]
              tma_load_2d(sa + c * 8192, &tmap_a, &sb->full[stage], mt * G_BM + c * 64, kb * G_BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d(sbm, &tmap_b, &sb->full[stage], kb * G_BK, nt * G_BN);           // box [64 k, 256 n]
          } else {
#pragma unroll
            for (int c = 0; c < G_BN / 64; ++c)                                            // box [64 n, 64 k]
              tma_load_2d(sbm + c * 8192, &tmap_b, &sb->full[stage], nt * G_BN + c * 64, kb * G_BK);
          }
          if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop, single elected lane issues (see head_pair.cuh)
    constexpr uint32_t idesc = make_idesc(G_BM, G_BN, A_MN, B_MN);
    constexpr uint32_t HI = desc_hi32(1024);
    constexpr uint32_t A_LOF = desc_lo_flags(A_MN ? 8192 : 16);
    constexpr uint32_t B_LOF = desc_lo_flags(B_MN ? 8192 : 16);
    constexpr uint32_t a_step = (A_MN ? 2048 : 32) >> 4;
    constexpr uint32_t b_step = (B_MN ? 2048 : 32) >> 4;
    const uint32_t smem_base = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = worker; item < total_items; item += num_workers) {
      int mt, nt, kb0, kb1;
      decode(item, mt, nt, kb0, kb1);
      mbar_wait(&sb->tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + acc * G_BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&sb->full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a = ((smem_base + stage * G_STAGE_BYTES) >> 4);
          const uint32_t b = a + (G_A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < G_BK / 16; ++k)
            umma_bf16(d, desc64((a + k * a_step) | A_LOF, HI), desc64((b + k * b_step) | B_LOF, HI), idesc,
                      (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&sb->empty[stage]);
          if (kb == kb1 - 1) umma_commit(&sb->tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
      }
      if (kb1 <= kb0) {                      // empty split: nothing was issued, still hand the stage to the epilogue
        if (elect_one()) umma_commit(&sb->tmem_full[acc]);
        __syncwarp();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;   // column half of the 256-wide tile
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = worker; item < total_items; item += num_workers) {
      int mt, nt, kb0, kb1;
      decode(item, mt, nt, kb0, kb1);
      const int m = mt * G_BM + quad * 32 + lane;
      long long orow = m;
      bool row_ok = m < p.M;
      if (OUT == OUT_RED_F32 && p.row_map != nullptr && row_ok) {
        const int r = __ldg(p.row_map + m);
        row_ok = r >= 0;
        orow = r;
      }
      mbar_wait(&sb->tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t0 = tmem_base + (uint32_t(quad * 32) << 16) + acc * G_BN + half * 128;
      const bool empty_k = kb1 <= kb0;   // split with no k-blocks: accumulator is stale, contributes nothing
      if constexpr (OUT == OUT_BF16) {
        // registers -> 128B-swizzled smem boxes -> TMA bulk store (full 128-byte lines; a direct store from the
        // TMEM register layout writes 16 bytes per lane to 32 different rows and is LSU-bound)
        if (warp == 4 && lane == 0) tma_store_wait_read();          // previous tile has left the staging buffer
        named_bar_sync(1, 256);
        const int r = quad * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          uint32_t v[32];
          tmem_ld32(t0 + c, v);
          tmem_ld_wait();
          const int col = half * 128 + c;                   // first of 32 columns
          uint8_t* box = ostage + (col >> 6) * (G_BM * 128);
          const int ch0 = (col & 63) >> 3;                  // first 16-byte chunk inside the 128-byte row
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v[8 * i + 0]), __uint_as_float(v[8 * i + 1]));
            o.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3]));
            o.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5]));
            o.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7]));
            *reinterpret_cast<uint4*>(box + swz128(r, ch0 + i)) = o;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sb->tmem_empty[acc]);   // accumulator drained: MMA may reuse the stage
        fence_proxy_async();
        named_bar_sync(1, 256);
        if (warp == 4 && lane == 0 && !empty_k) {
#pragma unroll
          for (int b = 0; b < G_BN / 64; ++b)
            if (nt * G_BN + b * 64 < p.N) tma_store_2d(&tmap_o, ostage + b * (G_BM * 128), nt * G_BN + b * 64, mt * G_BM);
          tma_store_commit();
        }
      } else {
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint32_t r[32];
        tmem_ld32(t0 + c, r);
        tmem_ld_wait();
        const int n0 = nt * G_BN + half * 128 + c;
        if (row_ok && !empty_k) {
          if constexpr (OUT == OUT_F32) {
            float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + n0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (n0 + i < p.N)
                *reinterpret_cast<uint4*>(o + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
            }
          } else {
            float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + n0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (n0 + i < p.N)
                red_add_v4(o + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]),
                           __uint_as_float(r[i + 3]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sb->tmem_empty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if constexpr (OUT == OUT_BF16) {
      if (warp == 4 && lane == 0) tma_store_wait_all();     // global writes complete before the kernel ends
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace hc
