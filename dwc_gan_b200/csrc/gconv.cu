// gconv: convolution forward / data-gradient as a multi-tap shifted GEMM over haloed NHWC buffers.
//
//  * gconv_tc_kernel   : tcgen05 (UMMA) implicit GEMM.  A tiles (128 pixels x 64 channels) and weight tiles
//                        (BN x 64) are staged by TMA (128B swizzle) through a 4-stage mbarrier ring, the
//                        fp32 accumulator lives in TMEM, the epilogue reads it back with tcgen05.ld.
//  * gconv_simt_kernel : CUDA-core version of the same abstract operation (fp32 validation mode, the
//                        3-channel image layers and the 4-channel head gradients).
#include "gconv.cuh"
#include <stdlib.h>

// =====================================================================================================
// SIMT kernel: 128 x 64 tile, 256 threads, each thread 8 rows x 4 columns, fp32 accumulate
// =====================================================================================================
constexpr int S_BM = 128, S_BN = 64, S_BK = 16;

template <typename T>
__global__ void __launch_bounds__(256) gconv_simt_kernel(const __grid_constant__ GConvDev p) {
  __shared__ float As[S_BK][S_BM + 4];
  __shared__ float Bs[S_BK][S_BN + 4];
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int col0 = blockIdx.y * S_BN;
  const T* __restrict__ A = reinterpret_cast<const T*>(p.a);
  const T* __restrict__ W = reinterpret_cast<const T*>(p.w);

  // loader roles
  const int lrow = tid >> 1;         // 0..127
  const int lk = (tid & 1) * 8;      // 0 or 8
  const RowCoord lrc = tile_row(p, tile, lrow);
  const int bcol = tid >> 2;         // 0..63
  const int bk = (tid & 3) * 4;      // 0,4,8,12
  // compute roles
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int C = p.C;
  for (int t = 0; t < p.ntaps; ++t) {
    const long long X = lrc.x + p.taps[t][0], Y = lrc.y + p.taps[t][1], Z = p.taps[t][2];
    const bool inb = X >= 0 && X < p.a_dim[1] && Y >= 0 && Y < p.a_dim[2] && lrc.n < p.a_dim[4];
    const long long abase = inb ? (lrc.n * p.a_str[4] + Z * p.a_str[3] + Y * p.a_str[2] + X * p.a_str[1]) : 0;
    for (int c0 = 0; c0 < C; c0 += S_BK) {
      // ---- stage A chunk
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int c = c0 + lk + j;
        float v = 0.f;
        if (inb && c < C) v = to_f<T>(A[abase + c]);
        As[lk + j][lrow] = v;
      }
      // ---- stage B chunk
      {
        int col = col0 + bcol;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int c = c0 + bk + j;
          float v = 0.f;
          if (col < p.ncols_padded && c < C) v = to_f<T>(W[(long long)col * p.K + (long long)t * C + c]);
          Bs[bk + j][bcol] = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < S_BK; ++k) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    RowCoord rc = tile_row(p, tile, ty * 8 + i);
    long long off;
    if (!out_offset(p, rc, &off)) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = col0 + tx * 4 + j;
      if (col >= p.ncols) continue;
      float v = acc[i][j] + (p.bias ? p.bias[col] : 0.f);
      if (p.out_dtype == DWC_F32) {
        float* o = reinterpret_cast<float*>(p.out) + off + col;
        *o = p.accumulate ? (*o + v) : v;
      } else {
        bf16* o = reinterpret_cast<bf16*>(p.out) + off + col;
        *o = __float2bfloat16_rn(p.accumulate ? (__bfloat162float(*o) + v) : v);
      }
    }
  }
}

// =====================================================================================================
// tcgen05 kernel
// =====================================================================================================
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;   // bf16 elements = one 128-byte swizzle row
constexpr int TC_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two warps per TMEM lane quadrant)

template <int BN> struct TcCfg {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;   // 16 KB
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int B_BYTES_AL = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 3 : 4);
  static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES_AL) + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, (BN >= 256 ? 1 : 2))
    gconv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ GConvDev p) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * (Cfg::A_BYTES + Cfg::B_BYTES_AL));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int col0 = blockIdx.y * BN;
  // Split K (few tiles, long K: the 4x4 stride-2 layers on 16x16 .. 4x4 maps, where one CTA per tile is bound by what
  // ONE SM can pull out of L2): the ksplit CTAs of a cluster (1,1,ksplit) each accumulate a contiguous share of the
  // (tap, channel block) iterations in their own TMEM, park the fp32 partial tile in their idle pipeline shared
  // memory, and each finishes 1/ksplit of the tile's columns from everybody's partial sums (reduce-scatter through
  // distributed shared memory, added in rank order: deterministic).
  const int ks = p.ksplit;
  const int ph = blockIdx.z / ks;                  // phase: weights ph*ncols_padded rows down, output ph*phase_out_off on
  const int kr = blockIdx.z - ph * ks;             // == %cluster_ctarank
  const int cblocks = p.C / TC_BK;
  const int all_kb = p.debug == 3 ? 0 : p.ntaps * cblocks;
  const int kb0 = all_kb / ks * kr;
  const int num_kb = all_kb / ks;

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
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  pdl_prologue();                     // everything above is CTA-local; the previous kernel's data from here on
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int tx = tile % p.tiles_x;
      int t2 = tile / p.tiles_x;
      int ty = t2 % p.tiles_y;
      int tn = t2 / p.tiles_y;
      const int x0 = tx * p.box_x, y0 = ty * p.box_y, n0 = tn * p.box_n;
      int stage = 0;
      uint32_t phase = 0;
      int t = kb0 / cblocks, cb = kb0 - t * cblocks;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int cx = x0 + p.taps[t][0], cy = y0 + p.taps[t][1], cz = p.taps[t][2];
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES);
        tma_load_5d(sA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], cb * TC_BK, cx, cy, cz, n0);
        tma_load_2d(sB + stage * Cfg::B_BYTES_AL, &tmB, &full_bar[stage], t * p.C + cb * TC_BK,
                    col0 + ph * p.ncols_padded);
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
        if (++cb == cblocks) {
          cb = 0;
          ++t;
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TC_BM, BN < 16 ? 16 : BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
        const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES_AL);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          uint64_t da = umma_desc_sw128(a_addr + k * 32, 16, 1024);
          uint64_t db = umma_desc_sw128(b_addr + k * 32, 16, 1024);
          umma_bf16(tmem_base, da, db, idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[stage]);   // frees the smem slot once these MMAs retire
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tmem_full);
    }
  } else {
    // ================= epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const RowCoord rc = tile_row(p, tile, r);
    long long off = 0;
    const bool valid = out_offset(p, rc, &off);
    off += ph * p.phase_out_off;
    long long d2off[4] = {0, 0, 0, 0};
    const int d2n = p.out2 ? out2_dests(p, rc, valid, d2off) : 0;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    constexpr int CHUNK = BN >= 32 ? 32 : 16;
    const bool staged = BN >= 64 && p.out_dtype == DWC_BF16 && (p.ncols & 7) == 0 && col0 + BN <= p.ncols &&
                        p.debug == 0;
    if (BN >= 64 && ks > 1) {
      // ---- split K: park the fp32 partial tile, [column][row], in the idle stage buffers ...
      float* part = reinterpret_cast<float*>(smem);
      {
        const int cw0 = ((warp - 2) >> 2) * (BN / 2);
#pragma unroll 1
        for (int cc = 0; cc < BN / 2; cc += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cw0 + cc), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) part[(cw0 + cc + j) * TC_BM + r] = __uint_as_float(v[j]);
        }
      }
      cluster_arrive_release();                            // every CTA's partial tile is complete and visible
      cluster_wait_acquire();
      // ---- ... then rank kr finishes columns [kr * BN/ks, (kr+1) * BN/ks) of the tile: the partial sums of all ranks,
      // added in rank order (deterministic), through distributed shared memory (~20 B/clk per SM, hence a slice each
      // instead of everything into one CTA).  Thread = one row x half of the slice's columns.
      const int SL = BN / ks;                              // >= 16
      const int et = threadIdx.x - 64;
      const int row = et & (TC_BM - 1), chalf = et >> 7;
      const int nc = SL >> 1;                              // this thread's columns, a multiple of 8
      const int sc0 = kr * SL + chalf * nc;                // first of them, inside the tile
      const RowCoord rc2 = tile_row(p, tile, row);
      long long off2 = 0;
      const bool valid2 = out_offset(p, rc2, &off2);
      off2 += ph * p.phase_out_off;
      long long e2off[4] = {0, 0, 0, 0};
      const int e2n = p.out2 ? out2_dests(p, rc2, valid2, e2off) : 0;
      float2* sst = reinterpret_cast<float2*>(smem + BN * TC_BM * 4);      // [4 row quarters][SL]
      uint32_t peer[8];
#pragma unroll
      for (int pr = 0; pr < 8; ++pr) peer[pr] = cluster_map_shared(smem_u32(smem), pr < ks ? pr : 0);
#pragma unroll 1
      for (int g8 = 0; g8 < nc; g8 += 8) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = 0.f;
        const uint32_t eoff = ((sc0 + g8) * TC_BM + row) * 4;
#pragma unroll
        for (int pr = 0; pr < 8; ++pr) {
          if (pr < ks) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] += ld_cluster_f32(peer[pr] + eoff + e * TC_BM * 4);
          }
        }
        const int cg = col0 + sc0 + g8;                    // global output column
        if (p.bias) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + cg));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + cg + 4));
          f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
          f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
        }
        uint4 sv;
        if (valid2) {
          bf16* o = reinterpret_cast<bf16*>(p.out) + off2 + cg;
          if (p.accumulate) {
            float old[8];
            Vec8<bf16>::load(o, old);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] += old[e];
          }
          Vec8<bf16>::store(reinterpret_cast<bf16*>(&sv), f);
          *reinterpret_cast<uint4*>(o) = sv;
          if (p.out2) {
            const uint4 av = act8_bf16(sv, p.o2_act);
            bf16* o2 = reinterpret_cast<bf16*>(p.out2) + cg;
            for (int d = 0; d < e2n; ++d) *reinterpret_cast<uint4*>(o2 + e2off[d]) = av;
          }
        } else {
          Vec8<bf16>::store(reinterpret_cast<bf16*>(&sv), f);
        }
        if (p.stats) {
          // {sum, sum of squares} of the stored values over this warp's 32 rows, per column
          float rs[8];
          Vec8<bf16>::load(reinterpret_cast<const bf16*>(&sv), rs);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float a = valid2 ? rs[e] : 0.f, b = a * a;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              a += __shfl_xor_sync(0xffffffffu, a, o);
              b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            if (lane == 0) sst[((row >> 5) * SL) + chalf * nc + g8 + e] = make_float2(a, b);
          }
        }
      }
      if (p.stats) {
        epi_bar_sync();
        const int tpi = p.tiles_x * p.tiles_y, n_img = tile / tpi, t_img = tile - n_img * tpi;
        if (et < SL) {
          float2 t = make_float2(0.f, 0.f);
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) {
            const float2 u = sst[w4 * SL + et];
            t.x += u.x; t.y += u.y;
          }
          p.stats[((long long)n_img * tpi + t_img) * p.ncols + col0 + kr * SL + et] = t;
        }
      }
    } else if (staged) {
      // Coalesced epilogue.  A thread owns one accumulator row, so storing straight from registers puts the 32 lanes
      // of every store instruction on 32 different output rows (16 bytes each, >= 512 bytes apart).  Instead each
      // warp stages its 32 rows x BN/2 columns (bf16, XOR-swizzled 16-byte pieces) in the now idle pipeline shared
      // memory and writes whole row segments with consecutive lanes.  Two warps share a TMEM lane quadrant, each
      // taking half of the columns.
      constexpr int HB_ = BN / 2;                          // columns per warp
      constexpr int PIECES = HB_ / 8;                      // 16-byte pieces per row segment (4, 8 or 16)
      constexpr int RPI = 32 / PIECES;                     // rows per store instruction
      constexpr int ROWB = HB_ * 2;
      const int half = (warp - 2) >> 2;
      const int cw0 = half * HB_;                          // first column of this warp inside the tile
      uint8_t* stg = smem + (warp - 2) * (32 * ROWB);      // all MMAs have retired: the stage buffers are free
      // fused statistics: per-warp column sums go to shared memory behind the staging area, [8 warps][HB_] float2
      float2* sst = reinterpret_cast<float2*>(smem + 8 * 32 * ROWB);
#pragma unroll 1
      for (int cc = 0; cc < HB_; cc += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cw0 + cc), v);
        tmem_ld_wait();
        float fa[32];
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float* f = fa + j;
          if (p.bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + cw0 + cc + j));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + cw0 + cc + j + 4));
            f[0] = __uint_as_float(v[j]) + b0.x; f[1] = __uint_as_float(v[j + 1]) + b0.y;
            f[2] = __uint_as_float(v[j + 2]) + b0.z; f[3] = __uint_as_float(v[j + 3]) + b0.w;
            f[4] = __uint_as_float(v[j + 4]) + b1.x; f[5] = __uint_as_float(v[j + 5]) + b1.y;
            f[6] = __uint_as_float(v[j + 6]) + b1.z; f[7] = __uint_as_float(v[j + 7]) + b1.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j + e]);
          }
          const int piece = (cc + j) >> 3;
          Vec8<bf16>::store(reinterpret_cast<bf16*>(stg + lane * ROWB + ((piece ^ (lane & (PIECES - 1))) << 4)), f);
        }
        if (p.stats) sst[(warp - 2) * HB_ + cc + lane] = warp_col_stats32(fa, lane, valid);
      }
      if (p.stats) {
        // add the four row quadrants of each column half and write this tile's partial sums
        epi_bar_sync();
        const int tpi = p.tiles_x * p.tiles_y, n_img = tile / tpi, t_img = tile - n_img * tpi;
        for (int c = threadIdx.x - 64; c < BN; c += 256) {
          const int h2 = c / HB_, ccol = c - h2 * HB_;
          float2 t = make_float2(0.f, 0.f);
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) {
            const float2 u = sst[(h2 * 4 + w4) * HB_ + ccol];
            t.x += u.x; t.y += u.y;
          }
          p.stats[((long long)n_img * tpi + t_img) * p.ncols + col0 + c] = t;
        }
      }
      __syncwarp();
      bf16* outp = reinterpret_cast<bf16*>(p.out) + col0 + cw0;
#pragma unroll 4
      for (int i = 0; i < 32; i += RPI) {
        const int rl = i + lane / PIECES, piece = lane % PIECES;
        const long long roff = __shfl_sync(0xffffffffu, off, rl);
        const int rv = __shfl_sync(0xffffffffu, (int)valid, rl);
        const uint4 dv = *reinterpret_cast<const uint4*>(stg + rl * ROWB + ((piece ^ (rl & (PIECES - 1))) << 4));
        if (rv) {
          uint4* dst = reinterpret_cast<uint4*>(outp + roff + piece * 8);
          if (p.accumulate) {
            // out += result: the sum of two bf16 tensors (e.g. the skip-connection gradient already in `out`)
            float a[8], b[8];
            Vec8<bf16>::load(reinterpret_cast<const bf16*>(&dv), a);
            Vec8<bf16>::load(reinterpret_cast<const bf16*>(dst), b);
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] += b[e];
            Vec8<bf16>::store(reinterpret_cast<bf16*>(dst), a);
          } else {
            *dst = dv;
          }
        }
        if (p.out2) {
          // activated copy into the next block's reflect-haloed input buffer: the row's (up to 4) destinations
          const uint4 av = act8_bf16(dv, p.o2_act);
          bf16* o2 = reinterpret_cast<bf16*>(p.out2) + col0 + cw0 + piece * 8;
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            const long long doff = __shfl_sync(0xffffffffu, d2off[d], rl);
            const int dn = __shfl_sync(0xffffffffu, d2n, rl);
            if (d < dn) *reinterpret_cast<uint4*>(o2 + doff) = av;
          }
        }
      }
    } else if (warp < 6) {
#pragma unroll 1
    for (int cc = 0; cc < (p.debug == 2 ? 0 : BN); cc += CHUNK) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc;
      if (CHUNK == 32) tmem_ld32(taddr, v);
      else tmem_ld16(taddr, v);
      tmem_ld_wait();
      if (valid && p.debug != 1) {
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
  }
  if (ks > 1) {
    // first cluster barrier: partial tiles written (the epilogue warps passed it before reading them); second: every
    // rank has read its slice, the peers' shared memory may go away
    if (warp < 2) {
      cluster_arrive_release();
      cluster_wait_acquire();
    }
    cluster_arrive_release();
    cluster_wait_acquire();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// =====================================================================================================
// persistent tcgen05 kernel: one CTA per SM walks over (tile, column block) work items; the accumulator is
// double-buffered in TMEM so that the epilogue of item i overlaps the main loop of item i+1 and the per-CTA
// prologue (barrier init, TMEM allocation, descriptor prefetch) is paid once.  Used for multi-wave launches.
// =====================================================================================================
template <int BN> struct TcpCfg {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int B_BYTES_AL = (B_BYTES + 1023) / 1024 * 1024;
  // BN = 256: one CTA per SM (192 KB of stages, all 512 TMEM columns); narrower tiles: two CTAs per SM, i.e. two
  // producer / MMA-issuer threads per SM as in the non-persistent kernel
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 3 : 4);
  static constexpr int CTAS_PER_SM = BN >= 256 ? 1 : 2;
  static constexpr int STAT_BYTES = BN >= 64 ? 2 * 8 * (BN / 2) * 8 : 0;     // 2 buffers x 8 warps x BN/2 columns x float2
  static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES_AL) + 1024 + 512 + STAT_BYTES;
  static constexpr int ACC_COLS = BN < 32 ? 32 : BN;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, (BN >= 256 ? 1 : 2))     // narrow tiles: two CTAs per SM (<= 102 registers)
    gconv_tcp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ GConvDev p) {
  using Cfg = TcpCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * (Cfg::A_BYTES + Cfg::B_BYTES_AL));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;        // [2]
  uint64_t* tmem_empty = tmem_full + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cblocks = p.C / TC_BK;
  const int num_kb = p.ntaps * cblocks;
  const int ntiles = p.tiles_x * p.tiles_y * p.tiles_n;
  const int ncb = (p.ncols_padded + BN - 1) / BN;
  const int per_phase = ntiles * ncb;                  // item = column block (fast) x tile x phase (slow)
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
      mbar_init(&tmem_empty[b], 8);                    // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  pdl_prologue();                     // everything above is CTA-local; the previous kernel's data from here on
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int ph = item / per_phase, it2 = item - ph * per_phase;
        const int tile = it2 / ncb, col0 = (it2 - tile * ncb) * BN + ph * p.ncols_padded;
        int tx = tile % p.tiles_x;
        int t2 = tile / p.tiles_x;
        int ty = t2 % p.tiles_y;
        int tn = t2 / p.tiles_y;
        const int x0 = tx * p.box_x, y0 = ty * p.box_y, n0 = tn * p.box_n;
        for (int t = 0; t < p.ntaps; ++t) {
          const int cx = x0 + p.taps[t][0], cy = y0 + p.taps[t][1], cz = p.taps[t][2];
          for (int cb = 0; cb < cblocks; ++cb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES);
            tma_load_5d(sA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], cb * TC_BK, cx, cy, cz, n0);
            tma_load_2d(sB + stage * Cfg::B_BYTES_AL, &tmB, &full_bar[stage], t * p.C + cb * TC_BK, col0);
            if (++stage == Cfg::STAGES) {
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
      constexpr uint32_t idesc = umma_idesc_bf16(TC_BM, BN < 16 ? 16 : BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int li = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++li) {
        const int acc = li & 1;
        mbar_wait(&tmem_empty[acc], ((li >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_addr = tmem_base + acc * Cfg::ACC_COLS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES_AL);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            uint64_t da = umma_desc_sw128(a_addr + k * 32, 16, 1024);
            uint64_t db = umma_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16(d_addr, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ================= epilogue: warps 2..9; TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4 =========
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int HCOLS = Cfg::ACC_COLS / 2;             // columns per warp
    constexpr int CHUNK = HCOLS >= 32 ? 32 : 16;
    const int r = q * 32 + lane;
    int li = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++li) {
      const int ph = item / per_phase, it2 = item - ph * per_phase;
      const int tile = it2 / ncb, col0 = (it2 - tile * ncb) * BN;
      const int acc = li & 1;
      const RowCoord rc = tile_row(p, tile, r);
      long long off = 0;
      const bool valid = out_offset(p, rc, &off);
      off += ph * p.phase_out_off;
      long long d2off[4] = {0, 0, 0, 0};
      const int d2n = p.out2 ? out2_dests(p, rc, valid, d2off) : 0;
      mbar_wait(&tmem_full[acc], (li >> 1) & 1);
      tc_fence_after();
      // fused statistics (BN >= 64): per-warp column sums in shared memory, double-buffered over items
      float2* sst = reinterpret_cast<float2*>(smem + Cfg::STAGES * (Cfg::A_BYTES + Cfg::B_BYTES_AL) + 512) +
                    (li & 1) * 8 * HCOLS;
#pragma unroll 1
      for (int cc = 0; cc < HCOLS; cc += CHUNK) {
        uint32_t v[32];
        __syncwarp();                                    // reconverge after the divergent store path below
        const int ccol = half * HCOLS + cc;              // column inside the tile
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * Cfg::ACC_COLS + ccol);
        if (CHUNK == 32) tmem_ld32(taddr, v);
        else tmem_ld16(taddr, v);
        tmem_ld_wait();
        if (cc + CHUNK >= HCOLS) {
          // last TMEM read of this warp: hand the accumulator back to the MMA warp before the global stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (BN >= 64 && CHUNK == 32 && p.stats) {
          // the bias is added here once (the store path below then sees biased values): v <- v + bias, in place
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + ccol + j));
              v[j] = __float_as_uint(__uint_as_float(v[j]) + b.x);
              v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + b.y);
              v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + b.z);
              v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + b.w);
            }
          }
          sst[(warp - 2) * HCOLS + cc + lane] = warp_col_stats32(reinterpret_cast<const float*>(v), lane, valid);
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
              if (p.bias && !(BN >= 64 && p.stats)) {
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
              uint4 sv;
              Vec8<bf16>::store(reinterpret_cast<bf16*>(&sv), f);
              *reinterpret_cast<uint4*>(o + j) = sv;
              if (p.out2) {                              // activated copy into the next block's reflect-haloed input
                const uint4 av = act8_bf16(sv, p.o2_act);
                bf16* o2 = reinterpret_cast<bf16*>(p.out2) + cbase + j;
                for (int d = 0; d < d2n; ++d) *reinterpret_cast<uint4*>(o2 + d2off[d]) = av;
              }
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
      if (BN >= 64 && p.stats) {
        // add the four row quadrants of each column half and write this tile's partial sums (one barrier per item:
        // the buffer of item li is next written at item li + 2, after every warp has passed the barrier of li + 1)
        __syncwarp();
        epi_bar_sync();
        const int tpi = p.tiles_x * p.tiles_y, n_img = tile / tpi, t_img = tile - n_img * tpi;
        for (int c = threadIdx.x - 64; c < BN; c += 256) {
          const int h2 = c / HCOLS, cq = c - h2 * HCOLS;
          float2 t = make_float2(0.f, 0.f);
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) {
            const float2 u = sst[(h2 * 4 + w4) * HCOLS + cq];
            t.x += u.x; t.y += u.y;
          }
          p.stats[((long long)n_img * tpi + t_img) * p.ncols + col0 + c] = t;
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

template <int BN>
static int launch_tcp(const dwc_gconv_t* g, const GConvDev& d, int nitems, cudaStream_t st) {
  using Cfg = TcpCfg<BN>;
  CUtensorMap tmA, tmB;
  if (dwc_make_tmap5(&tmA, g->a, g->a_dim, g->a_str, g->box[0], g->box[1], 1, g->box[2])) return 1;
  if (dwc_make_tmap2(&tmB, g->w, (int64_t)g->ncols_padded * d.nphase, d.K, d.K, BN, TC_BK)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    DWC_CUDA(cudaFuncSetAttribute(gconv_tcp_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  const int slots = dwc_num_sms() * Cfg::CTAS_PER_SM;
  const int grid = nitems < slots ? nitems : slots;
  DWC_CUDA(dwc_launch_pdl(gconv_tcp_kernel<BN>, dim3(grid), dim3(TC_THREADS), Cfg::SMEM, st, 1, tmA, tmB, d));
  DWC_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// host entry
// =====================================================================================================
template <int BN>
static int launch_tc(const dwc_gconv_t* g, const GConvDev& d, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  CUtensorMap tmA, tmB;
  if (dwc_make_tmap5(&tmA, g->a, g->a_dim, g->a_str, g->box[0], g->box[1], 1, g->box[2])) return 1;
  if (dwc_make_tmap2(&tmB, g->w, (int64_t)g->ncols_padded * d.nphase, d.K, d.K, BN, TC_BK)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    DWC_CUDA(cudaFuncSetAttribute(gconv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  // split K over a cluster when the launch fills only a fraction of the SMs (see the kernel)
  GConvDev dd = d;
  dd.ksplit = 1;
  static int ksmax = -1;
  if (ksmax < 0) {
    const char* e = getenv("DWC_KSPLIT");
    ksmax = e ? atoi(e) : 8;
  }
  const int ctas = d.tiles_x * d.tiles_y * d.tiles_n * cdiv(g->ncols_padded, BN) * d.nphase;
  const int all_kb = d.ntaps * (d.C / TC_BK);
  dim3 grid(d.tiles_x * d.tiles_y * d.tiles_n, cdiv(g->ncols_padded, BN), d.nphase);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 1;
  attr.val.clusterDim.y = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  if (BN >= 64 && d.out_dtype == DWC_BF16 && g->ncols % BN == 0 && d.debug == 0 && all_kb >= 32) {
    // a split pays when the K loop it removes (~0.25 us per iteration of one CTA) outweighs the reduction (~5 us):
    // at least 8 iterations left per CTA, and all clusters co-resident (one wave)
    static int max_clusters[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = ksmax; k >= 2; k >>= 1) {
      if (k > 8 || all_kb % k || all_kb / k < 8 || BN / k < 16 || ctas * k > dwc_num_sms()) continue;
      if (max_clusters[k] == 0) {
        cfg.gridDim = dim3(k * 16, 1, 1);
        attr.val.clusterDim.z = 1;
        attr.val.clusterDim.x = k;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, gconv_tc_kernel<BN>, &cfg) != cudaSuccess) {
          cudaGetLastError();
          nc = -1;
        }
        attr.val.clusterDim.x = 1;
        max_clusters[k] = nc > 0 ? nc : -1;
      }
      if (ctas <= max_clusters[k]) {
        dd.ksplit = k;
        break;
      }
    }
  }
  grid.z = d.nphase * dd.ksplit;
  DWC_CUDA(dwc_launch_pdl(gconv_tc_kernel<BN>, grid, dim3(TC_THREADS), Cfg::SMEM, st, dd.ksplit, tmA, tmB, dd));
  DWC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dwc_gconv(const dwc_gconv_t* g, dwc_stream_t stream) {
  DWC_CHECK(g != nullptr, "dwc_gconv: null params");
  DWC_CHECK(g->ntaps > 0 && g->ntaps <= DWC_MAX_TAPS, "dwc_gconv: ntaps %d out of range", g->ntaps);
  const bool halo = g->backend == DWC_TC_HALO || g->backend == DWC_TC_HALO1;
  DWC_CHECK(halo || g->box[0] * g->box[1] * g->box[2] == 128, "dwc_gconv: box must cover 128 rows");
  DWC_CHECK(g->a_str[0] == 1, "dwc_gconv: channel stride must be 1");
  const int nphase = g->nphase > 1 ? g->nphase : 1;
  if (nphase > 1) {
    const bool tc_ok = g->backend == DWC_TC && getenv("DWC_CG2") == nullptr &&
                       g->phase_w_off == (int64_t)g->ncols_padded * g->ntaps * g->a_dim[0];
    if (!tc_ok) {
      // kernels without phase support: one launch per phase
      const int64_t wes = g->dtype == DWC_BF16 ? 2 : 4, oes = g->out_dtype == DWC_BF16 ? 2 : 4;
      for (int ph = 0; ph < nphase; ++ph) {
        dwc_gconv_t q = *g;
        q.nphase = 1;
        q.w = reinterpret_cast<const char*>(g->w) + ph * g->phase_w_off * wes;
        q.out = reinterpret_cast<char*>(g->out) + ph * g->phase_out_off * oes;
        if (dwc_gconv(&q, stream)) return 1;
      }
      return 0;
    }
  }
  GConvDev d;
  memset(&d, 0, sizeof(d));
  d.a = g->a; d.w = g->w; d.bias = g->bias; d.out = g->out;
  for (int i = 0; i < 5; ++i) { d.a_dim[i] = g->a_dim[i]; d.a_str[i] = g->a_str[i]; }
  for (int i = 0; i < 3; ++i) d.o_str[i] = g->o_str[i];
  d.box_x = g->box[0]; d.box_y = g->box[1]; d.box_n = g->box[2];
  d.tiles_x = g->tiles[0]; d.tiles_y = g->tiles[1]; d.tiles_n = g->tiles[2];
  d.valid_x = g->valid[0]; d.valid_y = g->valid[1]; d.valid_n = g->valid[2];
  d.flat = g->flat; d.flat_img = g->flat_img; d.flat_pitch = g->flat_pitch; d.flat_h = g->flat_h; d.flat_w = g->flat_w;
  d.ntaps = g->ntaps; d.C = (int)g->a_dim[0]; d.K = g->ntaps * d.C;
  d.ncols = g->ncols; d.ncols_padded = g->ncols_padded;
  d.out_dtype = g->out_dtype; d.accumulate = g->accumulate;
  d.nphase = nphase; d.phase_out_off = g->phase_out_off;
  d.stats = reinterpret_cast<float2*>(g->stats);
  d.out2 = g->out2; d.o2_halo = g->out2_halo; d.o2_layout = g->out2_layout; d.o2_act = g->out2_act;
  if (g->out2) {
    const int npp = g->ncols_padded;
    const int bn = npp % 256 == 0 ? 256 : (npp % 128 == 0 ? 128 : 64);
    DWC_CHECK(g->backend == DWC_TC && g->dtype == DWC_BF16 && g->out_dtype == DWC_BF16 && !g->flat && nphase == 1 &&
                  g->ncols % 64 == 0 && g->ncols == g->ncols_padded && g->ncols % bn == 0 && !g->accumulate &&
                  (g->out2_act == 1 || g->out2_act == 2) && g->out2_halo >= 0 &&
                  (g->out2_halo == 0 || (g->valid[0] >= 2 * g->out2_halo + 2 && g->valid[1] >= 2 * g->out2_halo + 2)) &&
                  (g->out2_layout == 0 || (((g->valid[0] + 2 * g->out2_halo) & 1) == 0 && ((g->valid[1] + 2 * g->out2_halo) & 1) == 0)),
              "dwc_gconv: the activated second output needs the tcgen05 backend, bf16, ncols %% 64 == 0 and a haloed "
              "destination with at most one mirror image per axis");
  }
  if (g->stats) {
    DWC_CHECK(g->backend == DWC_TC && g->dtype == DWC_BF16 && g->out_dtype == DWC_BF16 && !g->flat && nphase == 1 &&
                  g->box[2] == 1 && g->ncols % 64 == 0 && g->ncols == g->ncols_padded && !g->accumulate,
              "dwc_gconv: fused statistics need the tcgen05 backend, bf16 output, one image per tile and ncols %% 64 == 0");
  }
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("DWC_GCONV_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    d.debug = dbg;
  }
  for (int t = 0; t < g->ntaps; ++t)
    for (int j = 0; j < 3; ++j) d.taps[t][j] = g->taps[t * 3 + j];
  cudaStream_t st = as_stream(stream);
  const long long ntiles = (long long)d.tiles_x * d.tiles_y * d.tiles_n;
  DWC_CHECK(ntiles > 0 && ntiles < 2147483647LL, "dwc_gconv: bad tile count");

  if (g->backend == DWC_TC || halo) {
    DWC_CHECK(g->dtype == DWC_BF16, "dwc_gconv: tcgen05 backend needs bf16 operands");
    DWC_CHECK(d.C % TC_BK == 0, "dwc_gconv: tcgen05 backend needs C %% 64 == 0 (C=%d)", d.C);
#ifdef DWC_EXPERIMENTAL
    if (halo) {
      int ks = 1;
      while (ks * ks < g->ntaps) ++ks;
      DWC_CHECK(ks * ks == g->ntaps, "dwc_gconv(halo): ntaps %d is not a square window", g->ntaps);
      return dwc_launch_gconv_halo(g, d, ks, st);
    }
    static int cg2 = -1;
    if (cg2 < 0) {
      const char* e = getenv("DWC_CG2");
      cg2 = e ? atoi(e) : 0;
    }
    if (cg2 && ntiles >= 2 && !g->stats && !g->out2) {
      const int rc = dwc_launch_gconv_tc2(g, d, st);
      if (rc >= 0) return rc;
    }
#else
    DWC_CHECK(!halo && getenv("DWC_CG2") == nullptr,
              "dwc_gconv: the halo-tile / CTA-pair kernels are parked (csrc/experimental); build with DWC_EXPERIMENTAL=1");
#endif
    {
      static int persist = -1;
      if (persist < 0) {
        const char* e = getenv("DWC_PERSISTENT");
        persist = e ? atoi(e) : 1;
      }
      const int npp = g->ncols_padded;
      const int bnp = npp % 256 == 0 ? 256 : (npp % 128 == 0 ? 128 : (npp % 64 == 0 ? 64 : (npp == 16 ? 16 : 0)));
      const int nitems = bnp ? (int)ntiles * cdiv(npp, bnp) * d.nphase : 0;
      if (persist && bnp && nitems > dwc_num_sms() * (bnp >= 256 ? 1 : 2) && d.debug == 0) {   // more than one wave of CTAs
        if (bnp == 256) return launch_tcp<256>(g, d, nitems, st);
        if (bnp == 128) return launch_tcp<128>(g, d, nitems, st);
        if (bnp == 64) return launch_tcp<64>(g, d, nitems, st);
        return launch_tcp<16>(g, d, nitems, st);
      }
    }
    const int np = g->ncols_padded;
    static int bn256 = -1;
    if (bn256 < 0) {
      const char* e = getenv("DWC_BN256");
      bn256 = e ? atoi(e) : 1;
    }
    if (np % 256 == 0 && bn256) return launch_tc<256>(g, d, st);
    if (np % 128 == 0) return launch_tc<128>(g, d, st);
    if (np % 64 == 0) return launch_tc<64>(g, d, st);
    if (np == 16) return launch_tc<16>(g, d, st);
    DWC_CHECK(false, "dwc_gconv: unsupported padded column count %d for tcgen05", np);
  }
  dim3 grid((unsigned)ntiles, cdiv(g->ncols, S_BN));
  if (g->dtype == DWC_F32)
    gconv_simt_kernel<float><<<grid, 256, 0, st>>>(d);
  else
    gconv_simt_kernel<bf16><<<grid, 256, 0, st>>>(d);
  DWC_LAUNCH_CHECK();
  return 0;
}
