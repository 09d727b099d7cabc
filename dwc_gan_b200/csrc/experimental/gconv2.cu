// gconv_tc2: the tap-by-tap tcgen05 implicit GEMM of gconv.cu on CTA PAIRS (tcgen05 cta_group::2).
//
// Two CTAs of a 2-CTA cluster (the two SMs of a TPC) own two consecutive 128-pixel tiles.  One thread of the leader
// CTA issues tcgen05.mma.cta_group::2 with M = 256: each SM multiplies its own 128-row A tile with the full
// BN-column weight tile, of which each CTA stages only HALF (BN/2 rows) in its shared memory - the tensor cores
// read the other half from the peer.  Per SM that halves the weight bytes fetched from L2 and read from shared
// memory per MMA, which is what lifts the single-CTA ceiling (measured ~1.1 PFLOP/s, profiles/) on this part.
//
//  barriers : full[s] lives in the LEADER (both CTAs' TMA loads complete_tx on it, .cta_group::2 form),
//             empty[s] and tmem_full live in both CTAs and are signalled by multicast tcgen05.commit;
//  roles    : warp 0 TMA producer (both CTAs), warp 1 MMA issuer (leader only) + TMEM alloc/dealloc (both),
//             warps 2-5 epilogue (both: TMEM lanes 0..127 of each CTA hold its own tile).
#include "../gconv.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int T2_BM = 128;
constexpr int T2_BK = 64;
constexpr int T2_THREADS = 192;

template <int BN> struct Tc2Cfg {
  static constexpr int A_BYTES = T2_BM * T2_BK * 2;            // 16 KB: this CTA's pixel tile
  static constexpr int B_BYTES = (BN / 2) * T2_BK * 2;         // this CTA's half of the weight tile
  static constexpr int B_BYTES_AL = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGE = A_BYTES + B_BYTES_AL;
  static constexpr int STAGES = BN >= 256 ? 6 : (BN >= 128 ? 8 : 8);
  static constexpr int SMEM = STAGES * STAGE + 1024 + 256;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared-memory location of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void tma2_load_5d(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1,
                                             int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                             int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <int COLS> __device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"((uint32_t)COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"((uint32_t)COLS) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

template <int BN>
__global__ void __launch_bounds__(T2_THREADS)
    gconv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ GConvDev p) {
  using Cfg = Tc2Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int tile = blockIdx.x;                  // tiles (2i, 2i+1) form a pair; a padding tile lies outside the grid
  const int col0 = blockIdx.y * BN;
  const int cblocks = p.C / T2_BK;
  const int num_kb = p.ntaps * cblocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      int tx = tile % p.tiles_x;
      int t2 = tile / p.tiles_x;
      int ty = t2 % p.tiles_y;
      int tn = t2 / p.tiles_y;
      const int x0 = tx * p.box_x, y0 = ty * p.box_y, n0 = tn * p.box_n;
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < p.ntaps; ++t) {
        const int cx = x0 + p.taps[t][0], cy = y0 + p.taps[t][1], cz = p.taps[t][2];
        for (int cb = 0; cb < cblocks; ++cb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
          const uint32_t fb = mapa_u32(&full_bar[stage], 0);        // the leader's barrier
          uint8_t* s = smem + stage * Cfg::STAGE;
          tma2_load_5d(s, &tmA, fb, cb * T2_BK, cx, cy, cz, n0);
          tma2_load_2d(s + Cfg::A_BYTES, &tmB, fb, t * p.C + cb * T2_BK, col0 + (int)rank * (BN / 2));
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE);
        const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < T2_BK / 16; ++k) {
          uint64_t da = umma_desc_sw128(a_addr + k * 32, 16, 1024);
          uint64_t db = umma_desc_sw128(b_addr + k * 32, 16, 1024);
          umma2_bf16(tmem_base, da, db, idesc, (kb | k) != 0);
        }
        umma2_commit_mc(&empty_bar[stage], 3);       // frees the slot in both CTAs once these MMAs retire
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma2_commit_mc(tmem_full, 3);
    }
  } else {
    // ================= epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const RowCoord rc = tile_row(p, tile, r);
    long long off = 0;
    const bool valid = out_offset(p, rc, &off) && tile < p.tiles_x * p.tiles_y * p.tiles_n;   // not the padding CTA
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int cc = 0; cc < BN; cc += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
      tmem_ld_wait();
      if (valid) {
        const int cbase = col0 + cc;
        if (p.out_dtype == DWC_BF16) {
          bf16* o = reinterpret_cast<bf16*>(p.out) + off + cbase;
          if (cbase + 32 <= p.ncols && (p.ncols & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j + e]) + (p.bias ? __ldg(p.bias + cbase + j + e) : 0.f);
              if (p.accumulate) {
                float old[8];
                Vec8<bf16>::load(o + j, old);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] += old[e];
              }
              Vec8<bf16>::store(o + j, f);
            }
          } else {
            for (int j = 0; j < 32; ++j) {
              if (cbase + j < p.ncols) {
                float f = __uint_as_float(v[j]) + (p.bias ? p.bias[cbase + j] : 0.f);
                if (p.accumulate) f += __bfloat162float(o[j]);
                o[j] = __float2bfloat16_rn(f);
              }
            }
          }
        } else {
          float* o = reinterpret_cast<float*>(p.out) + off + cbase;
          for (int j = 0; j < 32; ++j) {
            if (cbase + j < p.ncols) {
              float f = __uint_as_float(v[j]) + (p.bias ? p.bias[cbase + j] : 0.f);
              if (p.accumulate) f += o[j];
              o[j] = f;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // neither CTA leaves (or frees TMEM) while the pair's MMAs / commits are in flight
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN>
int launch_tc2(const dwc_gconv_t* g, const GConvDev& d, cudaStream_t st) {
  using Cfg = Tc2Cfg<BN>;
  CUtensorMap tmA, tmB;
  if (dwc_make_tmap5(&tmA, g->a, g->a_dim, g->a_str, g->box[0], g->box[1], 1, g->box[2])) return 1;
  if (dwc_make_tmap2(&tmB, g->w, g->ncols_padded, d.K, d.K, BN / 2, T2_BK)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    DWC_CUDA(cudaFuncSetAttribute(gconv_tc2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  const int ntiles = d.tiles_x * d.tiles_y * d.tiles_n;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((ntiles + 1) / 2 * 2, cdiv(g->ncols_padded, BN), 1);
  cfg.blockDim = dim3(T2_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DWC_CUDA(cudaLaunchKernelEx(&cfg, gconv_tc2_kernel<BN>, tmA, tmB, d));
  return 0;
}


// =====================================================================================================
// Persistent CTA-pair kernel (opt-in, DWC_CG2=2; NOT yet measured): the pair mechanics above inside the persistent
// structure of gconv_tcp_kernel - one cluster per SM pair walks over (tile pair, column block, phase) items, the
// accumulators are double-buffered in TMEM (2 x BN columns per CTA) and the eight epilogue warps of each CTA drain item i
// while the leader issues the MMAs of item i+1.  Why: the single-CTA kernel is bound by L2 -> SM operand traffic
// (681 MB per G7 launch at the chip's ~6300 B/clk L2 cap, profiles/r01g_ncu_g7_b48.md); M = 256 per weight tile cuts
// the bytes per FLOP by a third.
//   full[s]       : leader only, both CTAs' TMA loads complete_tx on it
//   empty[s]      : both CTAs, multicast tcgen05.commit from the leader
//   tmem_full[b]  : both CTAs, multicast tcgen05.commit
//   tmem_empty[b] : leader only, 16 arrivals (8 epilogue warps x 2 CTAs, remote arrive from the peer)
// =====================================================================================================
constexpr int T2P_THREADS = 320;

template <int BN> struct Tc2pCfg {
  static constexpr int A_BYTES = T2_BM * T2_BK * 2;
  static constexpr int B_BYTES = (BN / 2) * T2_BK * 2;
  static constexpr int B_BYTES_AL = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGE = A_BYTES + B_BYTES_AL;
  static constexpr int STAGES = BN >= 256 ? 6 : 8;
  static constexpr int SMEM = STAGES * STAGE + 1024 + 512;
  static constexpr int ACC_COLS = BN < 32 ? 32 : BN;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
};

__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(T2P_THREADS)
    gconv_tc2p_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ GConvDev p) {
  using Cfg = Tc2pCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;        // [2]
  uint64_t* tmem_empty = tmem_full + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int cblocks = p.C / T2_BK;
  const int num_kb = p.ntaps * cblocks;
  const int ntiles = p.tiles_x * p.tiles_y * p.tiles_n;
  const int ntp = (ntiles + 1) >> 1;                    // tile pairs (the last one may hold a padding tile)
  const int ncb = (p.ncols_padded + BN - 1) / BN;
  const int per_phase = ntp * ncb;
  const int nitems = per_phase * p.nphase;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 16);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (both CTAs): own pixel tile + own half of the weight tile =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < nitems; item += nclusters) {
        const int ph = item / per_phase, it2 = item - ph * per_phase;
        const int tp = it2 / ncb, col0 = (it2 - tp * ncb) * BN + ph * p.ncols_padded;
        const int tile = 2 * tp + (int)rank;
        int tx = tile % p.tiles_x;
        int t2 = tile / p.tiles_x;
        int ty = t2 % p.tiles_y;
        int tn = t2 / p.tiles_y;                        // == tiles_n for the padding tile: the loads are zero-filled
        const int x0 = tx * p.box_x, y0 = ty * p.box_y, n0 = tn * p.box_n;
        for (int t = 0; t < p.ntaps; ++t) {
          const int cx = x0 + p.taps[t][0], cy = y0 + p.taps[t][1], cz = p.taps[t][2];
          for (int cb = 0; cb < cblocks; ++cb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
            const uint32_t fb = mapa_u32(&full_bar[stage], 0);
            uint8_t* s = smem + stage * Cfg::STAGE;
            tma2_load_5d(s, &tmA, fb, cb * T2_BK, cx, cy, cz, n0);
            tma2_load_2d(s + Cfg::A_BYTES, &tmB, fb, t * p.C + cb * T2_BK, col0 + (int)rank * (BN / 2));
            if (++stage == Cfg::STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int li = 0;
      for (int item = cluster_id; item < nitems; item += nclusters, ++li) {
        const int acc = li & 1;
        mbar_wait(&tmem_empty[acc], ((li >> 1) & 1) ^ 1);       // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_addr = tmem_base + acc * Cfg::ACC_COLS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < T2_BK / 16; ++k) {
            uint64_t da = umma_desc_sw128(a_addr + k * 32, 16, 1024);
            uint64_t db = umma_desc_sw128(b_addr + k * 32, 16, 1024);
            umma2_bf16(d_addr, da, db, idesc, (kb | k) != 0);
          }
          umma2_commit_mc(&empty_bar[stage], 3);
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma2_commit_mc(&tmem_full[acc], 3);
      }
    }
  } else {
    // ================= epilogue (both CTAs): warps 2..9; lane quadrant = warp % 4, column half = (warp - 2) / 4 ======
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int HCOLS = Cfg::ACC_COLS / 2;
    constexpr int CHUNK = HCOLS >= 32 ? 32 : 16;
    const int r = q * 32 + lane;
    int li = 0;
    for (int item = cluster_id; item < nitems; item += nclusters, ++li) {
      const int ph = item / per_phase, it2 = item - ph * per_phase;
      const int tp = it2 / ncb, col0 = (it2 - tp * ncb) * BN;
      const int tile = 2 * tp + (int)rank;
      const int acc = li & 1;
      const RowCoord rc = tile_row(p, tile, r);
      long long off = 0;
      const bool valid = out_offset(p, rc, &off) && tile < ntiles;
      off += ph * p.phase_out_off;
      mbar_wait(&tmem_full[acc], (li >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < HCOLS; cc += CHUNK) {
        uint32_t v[32];
        __syncwarp();
        const int ccol = half * HCOLS + cc;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * Cfg::ACC_COLS + ccol);
        if (CHUNK == 32) tmem_ld32(taddr, v);
        else tmem_ld16(taddr, v);
        tmem_ld_wait();
        if (cc + CHUNK >= HCOLS) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(&tmem_empty[acc], 0));
        }
        if (!valid) continue;
        const int cbase = col0 + ccol;
        if (p.out_dtype == DWC_BF16) {
          bf16* o = reinterpret_cast<bf16*>(p.out) + off + cbase;
          if (cbase + CHUNK <= p.ncols && (p.ncols & 7) == 0) {
#pragma unroll
            for (int j = 0; j < CHUNK; j += 8) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j + e]);
              if (p.bias) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + cbase + j));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + cbase + j + 4));
                f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
              }
              if (p.accumulate) {
                float old[8];
                Vec8<bf16>::load(o + j, old);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] += old[e];
              }
              Vec8<bf16>::store(o + j, f);
            }
          } else {
            for (int j = 0; j < CHUNK; ++j) {
              if (cbase + j < p.ncols) {
                float f = __uint_as_float(v[j]) + (p.bias ? p.bias[cbase + j] : 0.f);
                if (p.accumulate) f += __bfloat162float(o[j]);
                o[j] = __float2bfloat16_rn(f);
              }
            }
          }
        } else {
          float* o = reinterpret_cast<float*>(p.out) + off + cbase;
          for (int j = 0; j < CHUNK; ++j) {
            if (cbase + j < p.ncols) {
              float f = __uint_as_float(v[j]) + (p.bias ? p.bias[cbase + j] : 0.f);
              if (p.accumulate) f += o[j];
              o[j] = f;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN>
int launch_tc2p(const dwc_gconv_t* g, const GConvDev& d, cudaStream_t st) {
  using Cfg = Tc2pCfg<BN>;
  CUtensorMap tmA, tmB;
  if (dwc_make_tmap5(&tmA, g->a, g->a_dim, g->a_str, g->box[0], g->box[1], 1, g->box[2])) return 1;
  if (dwc_make_tmap2(&tmB, g->w, (int64_t)g->ncols_padded * d.nphase, d.K, d.K, BN / 2, T2_BK)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    DWC_CUDA(cudaFuncSetAttribute(gconv_tc2p_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  const int ntiles = d.tiles_x * d.tiles_y * d.tiles_n;
  const int nitems = ((ntiles + 1) / 2) * cdiv(g->ncols_padded, BN) * d.nphase;
  int clusters = dwc_num_sms() / 2;
  if (clusters > nitems) clusters = nitems;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * clusters, 1, 1);
  cfg.blockDim = dim3(T2P_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DWC_CUDA(cudaLaunchKernelEx(&cfg, gconv_tc2p_kernel<BN>, tmA, tmB, d));
  return 0;
}

}  // namespace

// returns -1 if the geometry is not handled by the pair kernel (the caller falls back to the single-CTA kernel)
int dwc_launch_gconv_tc2(const dwc_gconv_t* g, const GConvDev& d, cudaStream_t st) {
  const int np = g->ncols_padded;
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("DWC_CG2");
    mode = e ? atoi(e) : 0;
  }
  if (mode == 2) {                                     // persistent pair kernel (all phases of a stride-2 dgrad in one launch)
    if (np % 256 == 0) return launch_tc2p<256>(g, d, st);
    if (np % 128 == 0) return launch_tc2p<128>(g, d, st);
    if (np % 64 == 0) return launch_tc2p<64>(g, d, st);
    return -1;
  }
  if (d.nphase != 1) return -1;                        // the one-shot pair kernel has no phase dimension
  if (np % 256 == 0) return launch_tc2<256>(g, d, st);
  if (np % 128 == 0) return launch_tc2<128>(g, d, st);
  if (np % 64 == 0) return launch_tc2<64>(g, d, st);
  return -1;
}
