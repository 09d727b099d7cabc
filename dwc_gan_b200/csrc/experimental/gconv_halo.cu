// gconv_halo: stride-1 k x k convolution (forward over the reflect-haloed input, data gradient over the zero-haloed
// output gradient) as a tcgen05 implicit GEMM that fetches each input pixel k times instead of k*k times.
//
//  tile      : 8*MT (x) x 16 (y) output pixels of one image = MT accumulators of 128 rows x BN columns in TMEM;
//  A operand : per (64-channel slab, horizontal tap dx) ONE TMA box of 8*MT x (15+k) input pixels lands in shared
//              memory as [row][8*MT px][64 ch] (128-byte swizzle).  The k vertical taps dy of that column shift are the
//              same buffer viewed through UMMA descriptors whose start address advances by dy image rows (a multiple
//              of 1024 B, so every 8-row group stays aligned with the swizzle atom; sliding the window horizontally
//              inside one buffer also works - the swizzle is applied to absolute smem address bits - but such
//              unaligned starts were measured to halve the MMA issue rate, see tests/diag_halo.py and profiles/);
//  B operand : [BN][64] weight tile per (tap, slab) through an mbarrier ring, used by all MT accumulators: with
//              MT = 2 the weight bytes per MMA cycle halve, so the ring covers the TMA latency;
//  roles     : warp 0 weight producer, warp 1 MMA issuer, warp 2 lane 0 input producer, warps 2-5 epilogue
//              (tcgen05.ld -> +bias -> bf16/fp32 store).
#include "../gconv.cuh"

namespace {

constexpr int HT_THREADS = 192;
constexpr int HT_BK = 64;
constexpr int HT_NA = 3;          // input-tile stages (one per horizontal tap in flight)

template <int BN, int MT> struct HaloCfg {
  static constexpr int B_BYTES = BN * HT_BK * 2;
  static constexpr int B_BYTES_AL = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int NB = BN >= 256 ? 3 : (BN >= 128 ? 6 : 8);
  static constexpr int ACC_COLS = BN < 32 ? 32 : BN;          // TMEM columns per accumulator
  static constexpr int TMEM_COLS = ACC_COLS * MT < 32 ? 32 : ACC_COLS * MT;
  static constexpr int PITCH = 8 * MT;                         // pixels per staged image row
  static constexpr int ROW_BYTES = PITCH * 128;
};

// K-major SW128 descriptor with an explicit 8-row-group stride.  The start address may be any multiple of 128 bytes
// inside the TMA-written buffer: the hardware applies the 128B swizzle to the absolute shared-memory address bits
// (measured on B200: a pixel-shifted start with matrix-base-offset 0 reads exactly the rows TMA wrote; setting the
// base-offset field to the row phase does NOT - tests/diag_halo.py), so a shifted window needs no re-staging.
__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(16 >> 4) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN, int MT>
__global__ void __launch_bounds__(HT_THREADS)
    gconv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ GConvDev p, const int ksize) {
  using Cfg = HaloCfg<BN, MT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_stage = (15 + ksize) * Cfg::ROW_BYTES;
  uint8_t* sA = smem;
  uint8_t* sB = smem + HT_NA * a_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + Cfg::NB * Cfg::B_BYTES_AL);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + HT_NA;
  uint64_t* b_full = bars + 2 * HT_NA;
  uint64_t* b_empty = b_full + Cfg::NB;
  uint64_t* tmem_full = b_empty + Cfg::NB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int col0 = blockIdx.y * BN;
  const int cblocks = p.C / HT_BK;
  int tx = tile % p.tiles_x;
  int t2 = tile / p.tiles_x;
  const int x0 = tx * 8 * MT, y0 = (t2 % p.tiles_y) * 16, n0 = t2 / p.tiles_y;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < HT_NA; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < Cfg::NB; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= weight producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int cb = 0; cb < cblocks; ++cb) {
        for (int dx = 0; dx < ksize; ++dx) {
          for (int dy = 0; dy < ksize; ++dy) {
            const int t = dy * ksize + dx;
            mbar_wait(&b_empty[stage], phase ^ 1);
            mbar_expect_tx(&b_full[stage], Cfg::B_BYTES);
            uint8_t* dst = sB + stage * Cfg::B_BYTES_AL;
            tma_load_2d(dst, &tmB, &b_full[stage], t * p.C + cb * HT_BK, col0);
            if (++stage == Cfg::NB) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN < 16 ? 16 : BN, 0, 0);
      int bs = 0, as = 0;
      uint32_t bphase = 0, aphase = 0;
      uint32_t any = 0;
      for (int cb = 0; cb < cblocks; ++cb) {
        for (int dx = 0; dx < ksize; ++dx) {
          mbar_wait(&a_full[as], aphase);
          const uint32_t a_base = smem_u32(sA + as * a_stage);
          for (int dy = 0; dy < ksize; ++dy) {
            mbar_wait(&b_full[bs], bphase);
            tc_fence_after();
            const uint32_t a_addr = a_base + (uint32_t)dy * Cfg::ROW_BYTES;
            const uint32_t b_addr = smem_u32(sB + bs * Cfg::B_BYTES_AL);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
#pragma unroll
              for (int k = 0; k < HT_BK / 16; ++k) {
                const uint64_t da = umma_desc_sw128_sbo(a_addr + m * 1024 + k * 32, Cfg::ROW_BYTES);
                const uint64_t db = umma_desc_sw128(b_addr + k * 32, 16, 1024);
                umma_bf16(tmem_base + m * Cfg::ACC_COLS, da, db, idesc, any | (uint32_t)(k != 0));
              }
            }
            any = 1;
            umma_commit(&b_empty[bs]);
            if (++bs == Cfg::NB) {
              bs = 0;
              bphase ^= 1;
            }
          }
          umma_commit(&a_empty[as]);
          if (++as == HT_NA) {
            as = 0;
            aphase ^= 1;
          }
        }
      }
      umma_commit(tmem_full);
    }
  } else {
    if (warp == 2 && lane == 0) {
      // ================= input-tile producer =================
      int as = 0;
      uint32_t aphase = 0;
      for (int cb = 0; cb < cblocks; ++cb) {
        for (int dx = 0; dx < ksize; ++dx) {
          mbar_wait(&a_empty[as], aphase ^ 1);
          mbar_expect_tx(&a_full[as], (uint32_t)a_stage);
          tma_load_5d(sA + as * a_stage, &tmA, &a_full[as], cb * HT_BK, x0 + dx, y0, 0, n0);
          if (++as == HT_NA) {
            as = 0;
            aphase ^= 1;
          }
        }
      }
    }
    __syncwarp();
    // ================= epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    constexpr int CHUNK = BN >= 32 ? 32 : 16;
#pragma unroll 1
    for (int mc = 0; mc < MT * BN; mc += CHUNK) {
      const int m = mc / BN, cc = mc - m * BN;
      RowCoord rc;
      rc.x = x0 + m * 8 + (r & 7);
      rc.y = y0 + (r >> 3);
      rc.n = n0;
      long long off = 0;
      const bool valid = out_offset(p, rc, &off);
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * Cfg::ACC_COLS + cc);
      if (CHUNK == 32) tmem_ld32(taddr, v);
      else tmem_ld16(taddr, v);
      tmem_ld_wait();
      if (valid) {
        const int cbase = col0 + cc;
        if (p.out_dtype == DWC_BF16) {
          bf16* o = reinterpret_cast<bf16*>(p.out) + off + cbase;
          if (cbase + CHUNK <= p.ncols && (p.ncols & 7) == 0) {
#pragma unroll
            for (int j = 0; j < CHUNK; j += 8) {
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
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN, int MT>
int launch_halo(const dwc_gconv_t* g, const GConvDev& d, int ksize, cudaStream_t st) {
  using Cfg = HaloCfg<BN, MT>;
  CUtensorMap tmA, tmB;
  if (dwc_make_tmap5(&tmA, g->a, g->a_dim, g->a_str, Cfg::PITCH, 15 + ksize, 1, 1)) return 1;
  if (dwc_make_tmap2(&tmB, g->w, g->ncols_padded, d.K, d.K, BN, HT_BK)) return 1;
  const int a_stage = (15 + ksize) * Cfg::ROW_BYTES;
  const int smem = HT_NA * a_stage + Cfg::NB * Cfg::B_BYTES_AL + 1024 + 256;
  DWC_CHECK(smem <= 227 * 1024, "dwc_gconv(halo): shared memory %d exceeds the SM", smem);
  static int attr_smem = 0;
  if (smem > attr_smem) {
    DWC_CUDA(cudaFuncSetAttribute(gconv_halo_kernel<BN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  dim3 grid(d.tiles_x * d.tiles_y * d.tiles_n, cdiv(g->ncols_padded, BN));
  gconv_halo_kernel<BN, MT><<<grid, HT_THREADS, smem, st>>>(tmA, tmB, d, ksize);
  DWC_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int dwc_launch_gconv_halo(const dwc_gconv_t* g, const GConvDev& d, int ksize, cudaStream_t st) {
  DWC_CHECK(ksize >= 1 && ksize <= 9, "dwc_gconv(halo): window %d out of range", ksize);
  const int mt = g->box[0] / 8;
  DWC_CHECK((mt == 1 || mt == 2) && g->box[0] == 8 * mt && g->box[1] == 16 && g->box[2] == 1,
            "dwc_gconv(halo): tile must be 8 x 16 x 1 or 16 x 16 x 1");
  DWC_CHECK(!g->flat, "dwc_gconv(halo): flat addressing is not supported");
  for (int t = 0; t < g->ntaps; ++t)
    DWC_CHECK(g->taps[t * 3] == t % ksize && g->taps[t * 3 + 1] == t / ksize && g->taps[t * 3 + 2] == 0,
              "dwc_gconv(halo): taps must be the full k x k window in (ky, kx) order");
  const int np = g->ncols_padded;
  if (mt == 2) {
    // two accumulators per CTA: 128-column blocks keep the weight stage at 16 KB (more stages in flight)
    if (np % 256 == 0) return launch_halo<256, 2>(g, d, ksize, st);
    if (np % 128 == 0) return launch_halo<128, 2>(g, d, ksize, st);
    if (np % 64 == 0) return launch_halo<64, 2>(g, d, ksize, st);
    if (np == 16) return launch_halo<16, 2>(g, d, ksize, st);
  } else {
    if (np % 256 == 0) return launch_halo<256, 1>(g, d, ksize, st);
    if (np % 128 == 0) return launch_halo<128, 1>(g, d, ksize, st);
    if (np % 64 == 0) return launch_halo<64, 1>(g, d, ksize, st);
    if (np == 16) return launch_halo<16, 1>(g, d, ksize, st);
  }
  DWC_CHECK(false, "dwc_gconv(halo): unsupported padded column count %d", np);
}
