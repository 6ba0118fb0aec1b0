// cta_group::2 variant of the tcgen05 GEMM (see gemm_tc.cuh for the operand conventions).
//
// Why: what limits these GEMMs is operand DELIVERY into the SM, not the tensor pipe: a single SM ingests ~54.5 B/clk
// through TMA (tools/tma_bench.cu; the same for L2- and HBM-resident sources, for 18 or 148 active SMs), while
// back-to-back 128x256x16 MMAs on resident operands run at their nominal 128 cycles (tools/mma_bench.cu).  A 1-CTA
// 128 x 256 tile needs 48 KB per 64-element k-block (512 MMA cycles) = 880 cycles of ingest; a CTA pair running ONE
// M=256 x N=256 MMA needs each CTA's 128 rows of A and only HALF of B: 32 KB = 590 cycles, ~87 % of the tensor peak.
//
// Cluster of 2 CTAs = one 256 x 256 output tile; CTA r owns rows [128r, 128r+128) (accumulator in its own TMEM).
//   producer (warp 0, both CTAs): own A tile + own half of B; completion bytes go to the LEADER's `full` barrier
//   MMA (warp 1, leader only)   : tcgen05.mma.cta_group::2, commits multicast to both CTAs' `empty` / `tmem_full`
//   epilogue (warps 4-11, both) : own 128 x 256 accumulator; releases the stage on the LEADER's `tmem_empty`
#pragma once
#include "gemm_tc.cuh"

namespace hc {

constexpr int G2_A_BYTES = G_BM * G_BK * 2;          // 16 KB: this CTA's 128 rows
constexpr int G2_B_BYTES = (G_BN / 2) * G_BK * 2;    // 16 KB: this CTA's half of the 256 columns
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
template <int OUT> struct Gemm2Cfg {
  static constexpr int STAGES = (OUT == 0) ? 4 : 6;
  static constexpr int SMEM_BYTES = STAGES * G2_STAGE_BYTES + (OUT == 0 ? G_OUT_STAGE_BYTES : 0) + 1024 + 256;
};
constexpr int G2_MAX_STAGES = 6;

struct Gemm2Smem {
  uint64_t full[G2_MAX_STAGES];
  uint64_t empty[G2_MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

template <bool A_MN, bool B_MN, int OUT>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm2_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_o, const GemmParams p) {
  constexpr int STAGES = Gemm2Cfg<OUT>::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ostage = smem + STAGES * G2_STAGE_BYTES;
  Gemm2Smem* sb = reinterpret_cast<Gemm2Smem*>(ostage + (OUT == OUT_BF16 ? G_OUT_STAGE_BYTES : 0));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = int(cluster_ctarank());
  const bool leader = crank == 0;
  const int m2_tiles = (p.M + 2 * G_BM - 1) / (2 * G_BM);
  const int tiles_mn = m2_tiles * p.num_n_tiles;
  const int total_items = tiles_mn * p.splits;
  const int worker = blockIdx.x >> 1, num_workers = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&sb->full[i], 2);        // leader's expect_tx arrive + the peer producer's remote arrive
      mbar_init(&sb->empty[i], 1);       // leader's multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sb->tmem_full[i], 1);   // leader's multicast commit
      mbar_init(&sb->tmem_empty[i], 16); // 8 epilogue warps of each CTA (used in the leader only)
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2cta<512>(&sb->tmem_base);
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = sb->tmem_base;

  // k-block activity flags of A's row tile mt2 (8 per 64-bit word); nullptr = dense
  const bool sparse = p.kact != nullptr;
  auto kact_word = [&](int mt2, int kb) -> unsigned long long {
    return __ldg(reinterpret_cast<const unsigned long long*>(p.kact + (size_t)mt2 * p.kact_ld + (kb & ~7)));
  };
  auto any_active = [&](int mt2, int kb0, int kb1) -> bool {        // (epilogue) does the item run any k-block?
    if (!sparse) return kb1 > kb0;
    for (int kb = kb0 & ~7; kb < kb1; kb += 8) {
      unsigned long long w = kact_word(mt2, kb);
      if (kb < kb0) w &= ~0ull << (8 * (kb0 - kb));
      if (kb + 8 > kb1) w &= ~0ull >> (8 * (kb + 8 - kb1));
      if (w) return true;
    }
    return false;
  };
  auto decode = [&](int item, int& mt2, int& nt, int& kb0, int& kb1) {
    const int sp = item / tiles_mn;
    const int r = item - sp * tiles_mn;
    mt2 = r / p.num_n_tiles;
    nt = r - mt2 * p.num_n_tiles;
    kb0 = sp * p.k_blocks_per_split;
    kb1 = min(p.num_k_blocks, kb0 + p.k_blocks_per_split);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = worker; item < total_items; item += num_workers) {
        int mt2, nt, kb0, kb1;
        decode(item, mt2, nt, kb0, kb1);
        const int m0 = mt2 * 2 * G_BM + crank * G_BM;          // my rows
        const int n0 = nt * G_BN + crank * (G_BN / 2);         // my half of the columns
        // flag words are fetched one 8-block group ahead: a dependent global load in front of a k-block would stall the
        // one producer thread for longer than the operand ring holds
        unsigned long long fw = (sparse && kb0 < kb1) ? kact_word(mt2, kb0) : ~0ull;
        unsigned long long fw_next = (sparse && (kb0 | 7) + 1 < kb1) ? kact_word(mt2, (kb0 | 7) + 1) : 0ull;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (sparse) {
            if ((kb & 7) == 0 && kb != kb0) {
              fw = fw_next;
              if (kb + 8 < kb1) fw_next = kact_word(mt2, kb + 8);
            }
            if (((fw >> (8 * (kb & 7))) & 0xffull) == 0) continue;       // all-zero block of A: nothing to fetch
          }
          mbar_wait(&sb->empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * G2_STAGE_BYTES;
          uint8_t* sbm = sa + G2_A_BYTES;
          if (leader) mbar_arrive_expect_tx(&sb->full[stage], 2 * G2_STAGE_BYTES);
          else mbar_arrive_cluster(&sb->full[stage], 0);
          if constexpr (!A_MN) {
            tma_load_2d_2cta(sa, &tmap_a, &sb->full[stage], kb * G_BK, m0);                 // box [64 k, 128 m]
          } else {
#pragma unroll
            for (int c = 0; c < G_BM / 64; ++c)                                               // box [64 m, 64 k]
              tma_load_2d_2cta(sa + c * 8192, &tmap_a, &sb->full[stage], m0 + c * 64, kb * G_BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d_2cta(sbm, &tmap_b, &sb->full[stage], kb * G_BK, n0);                // box [64 k, 128 n]
          } else {
#pragma unroll
            for (int c = 0; c < G_BN / 128; ++c)                                              // box [64 n, 64 k]
              tma_load_2d_2cta(sbm + c * 8192, &tmap_b, &sb->full[stage], n0 + c * 64, kb * G_BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = make_idesc(2 * G_BM, G_BN, A_MN, B_MN);
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
        int mt2, nt, kb0, kb1;
        decode(item, mt2, nt, kb0, kb1);
        mbar_wait(&sb->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * G_BN;
        unsigned long long fw = (sparse && kb0 < kb1) ? kact_word(mt2, kb0) : ~0ull;
        unsigned long long fw_next = (sparse && (kb0 | 7) + 1 < kb1) ? kact_word(mt2, (kb0 | 7) + 1) : 0ull;
        bool started = false;              // the first executed k-block overwrites the accumulator
        for (int kb = kb0; kb < kb1; ++kb) {
          if (sparse) {
            if ((kb & 7) == 0 && kb != kb0) {
              fw = fw_next;
              if (kb + 8 < kb1) fw_next = kact_word(mt2, kb + 8);
            }
            if (((fw >> (8 * (kb & 7))) & 0xffull) == 0) continue;
          }
          mbar_wait(&sb->full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a = ((smem_base + stage * G2_STAGE_BYTES) >> 4);
            const uint32_t b = a + (G2_A_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < G_BK / 16; ++k)
              umma_bf16_2cta(d, desc64((a + k * a_step) | A_LOF, HI), desc64((b + k * b_step) | B_LOF, HI), idesc,
                             (started || k > 0) ? 1u : 0u);
            umma_commit_2cta(&sb->empty[stage], uint16_t(3));
          }
          __syncwarp();
          started = true;
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete (commit tracks every MMA issued so far; with no k-block executed it fires at once and the
        // epilogue treats the tile as zeros)
        if (elect_one()) umma_commit_2cta(&sb->tmem_full[acc], uint16_t(3));
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = worker; item < total_items; item += num_workers) {
      int mt2, nt, kb0, kb1;
      decode(item, mt2, nt, kb0, kb1);
      const int mrow0 = mt2 * 2 * G_BM + crank * G_BM;
      const int m = mrow0 + quad * 32 + lane;
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
      const bool empty_k = kb1 <= kb0;                 // split without k-blocks: contributes nothing, stores nothing
      const bool zero_tile = !empty_k && !any_active(mt2, kb0, kb1);      // every k-block skipped: the product is zero
      if constexpr (OUT == OUT_BF16) {
        if (warp == 4 && lane == 0) tma_store_wait_read();
        named_bar_sync(1, 256);
        const int r = quad * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          uint32_t v[32];
          if (!zero_tile) {
            tmem_ld32(t0 + c, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0u;
          }
          const int col = half * 128 + c;
          uint8_t* box = ostage + (col >> 6) * (G_BM * 128);
          const int ch0 = (col & 63) >> 3;
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
        if (lane == 0) mbar_arrive_cluster(&sb->tmem_empty[acc], 0);
        fence_proxy_async();
        named_bar_sync(1, 256);
        if (warp == 4 && lane == 0 && !empty_k && mrow0 < p.M) {
#pragma unroll
          for (int b = 0; b < G_BN / 64; ++b)
            if (nt * G_BN + b * 64 < p.N) tma_store_2d(&tmap_o, ostage + b * (G_BM * 128), nt * G_BN + b * 64, mrow0);
          tma_store_commit();
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          uint32_t r[32];
          if (!zero_tile) {
            tmem_ld32(t0 + c, r);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = 0u;
          }
          const int n0 = nt * G_BN + half * 128 + c;
          if (row_ok && !empty_k && !(zero_tile && OUT == OUT_RED_F32)) {
            float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + n0;
            if constexpr (OUT == OUT_F32) {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                if (n0 + i < p.N) *reinterpret_cast<uint4*>(o + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                if (n0 + i < p.N)
                  red_add_v4(o + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]),
                             __uint_as_float(r[i + 3]));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&sb->tmem_empty[acc], 0);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if constexpr (OUT == OUT_BF16) {
      if (warp == 4 && lane == 0) tma_store_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();            // nobody leaves while the pair's MMAs / remote arrives may still touch this CTA
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta<512>(tmem_base);
  }
}

}  // namespace hc
