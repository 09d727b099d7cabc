// wgrad: convolution weight gradient  dW[ia, tap, ib] = sum_pixels dY[pix, ia] * Xpad[pix + tap, ib]
//
//  * wgrad_tc_kernel   : tcgen05 GEMM whose K dimension is the pixel index.  Both operands are NHWC tiles
//                        (64 pixels x 64 channels, channel-contiguous) staged by TMA, i.e. "MN-major" UMMA
//                        operands; the 128 x BN fp32 accumulator of one (tap, channel-block) pair lives in
//                        TMEM while the CTA streams its share of the pixels (split-K).
//  * wgrad_simt_kernel : CUDA-core version (fp32 validation mode and the 3 / 4 channel layers).
//  Partial sums are written to a workspace and reduced in a fixed order (deterministic).
#include "common.cuh"
#include <stdlib.h>

struct WgradDev {
  const void* m_ptr;   // operand whose channels become accumulator rows
  const void* n_ptr;   // operand whose channels become accumulator columns
  long long m_dim[5], m_str[5], n_dim[5], n_str[5];
  int m_tap, n_tap;    // which operand the tap offsets apply to
  int cm, cn;
  int box_x, box_y, box_n;
  int tiles_x, tiles_y, tiles_n;
  int ntiles, splits, tiles_per_split;
  int ntaps;
  float* ws;           // [split][tap][cm][cn]
  int ngroup;          // taps per item on the N operand (1 = classic); > 1: N boxes are the SAME channels of `ngroup`
                       // taps, read at pixel - tap (the M operand is untapped and the tiles walk ITS pixel grid)
  int ntaps_total;     // real tap count (items cover ceil(ntaps_total / ngroup) groups)
  int debug;           // diagnostics (DWC_WGRAD_DEBUG): 1 = no MMAs (memory pipeline only), 2 = no loads after the first ring fill
  int grp_plus;        // grouped mode, which side carries the taps: 0 = the N operand is read at pixel - tap (operands
                       // swapped, M = untapped padded input); 1 = the N operand IS the tapped input, read at pixel + tap
                       // (incl. its parity plane), M = untapped dY and the tiles walk dY's grid
  int taps[DWC_MAX_TAPS][3];
};

__device__ __forceinline__ void wg_tile_origin(const WgradDev& p, int tile, int* x0, int* y0, int* n0) {
  int tx = tile % p.tiles_x;
  int t2 = tile / p.tiles_x;
  *x0 = tx * p.box_x;
  *y0 = (t2 % p.tiles_y) * p.box_y;
  *n0 = (t2 / p.tiles_y) * p.box_n;
}

// =====================================================================================================
// SIMT: 64 x 64 output tile per CTA, 256 threads x (4 x 4), K = pixels in chunks of 16
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const __grid_constant__ WgradDev p) {
  __shared__ float Ms[16][64 + 4];
  __shared__ float Ns[16][64 + 4];
  const int tid = threadIdx.x;
  const int mblocks = (p.cm + 63) / 64, nblocks = (p.cn + 63) / 64;
  int wid = blockIdx.x;
  const int mb = wid % mblocks; wid /= mblocks;
  const int nb = wid % nblocks; wid /= nblocks;
  const int t = wid;
  const int split = blockIdx.y;
  const T* __restrict__ Mp = reinterpret_cast<const T*>(p.m_ptr);
  const T* __restrict__ Np = reinterpret_cast<const T*>(p.n_ptr);
  const int ty = tid >> 4, tx = tid & 15;
  const int lpix = tid >> 4;          // 0..15 pixel within chunk
  const int lch = (tid & 15) * 4;     // 4 channels
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int mdx = p.m_tap ? p.taps[t][0] : 0, mdy = p.m_tap ? p.taps[t][1] : 0, mdz = p.m_tap ? p.taps[t][2] : 0;
  const int ndx = p.n_tap ? p.taps[t][0] : 0, ndy = p.n_tap ? p.taps[t][1] : 0, ndz = p.n_tap ? p.taps[t][2] : 0;
  const int tile_begin = split * p.tiles_per_split;
  const int tile_end = min(p.ntiles, tile_begin + p.tiles_per_split);
  const int rows = p.box_x * p.box_y * p.box_n;
  for (int tile = tile_begin; tile < tile_end; ++tile) {
    int x0, y0, n0;
    wg_tile_origin(p, tile, &x0, &y0, &n0);
    for (int r0 = 0; r0 < rows; r0 += 16) {
      int r = r0 + lpix;
      int x = x0 + r % p.box_x;
      int r2 = r / p.box_x;
      int y = y0 + r2 % p.box_y;
      int n = n0 + r2 / p.box_y;
      {
        long long X = x + mdx, Y = y + mdy;
        bool inb = r < rows && X >= 0 && X < p.m_dim[1] && Y >= 0 && Y < p.m_dim[2] && n < p.m_dim[4];
        long long base = n * p.m_str[4] + (long long)mdz * p.m_str[3] + Y * p.m_str[2] + X * p.m_str[1];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int ch = mb * 64 + lch + j;
          Ms[lpix][lch + j] = (inb && ch < p.cm) ? to_f<T>(Mp[base + ch]) : 0.f;
        }
      }
      {
        long long X = x + ndx, Y = y + ndy;
        bool inb = r < rows && X >= 0 && X < p.n_dim[1] && Y >= 0 && Y < p.n_dim[2] && n < p.n_dim[4];
        long long base = n * p.n_str[4] + (long long)ndz * p.n_str[3] + Y * p.n_str[2] + X * p.n_str[1];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int ch = nb * 64 + lch + j;
          Ns[lpix][lch + j] = (inb && ch < p.cn) ? to_f<T>(Np[base + ch]) : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float4 a = *reinterpret_cast<const float4*>(&Ms[k][ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Ns[k][tx * 4]);
        float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  float* ws = p.ws + ((long long)split * p.ntaps + t) * p.cm * p.cn;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = mb * 64 + ty * 4 + i;
    if (m >= p.cm) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = nb * 64 + tx * 4 + j;
      if (n < p.cn) ws[(long long)m * p.cn + n] = acc[i][j];
    }
  }
}

// =====================================================================================================
// tcgen05: M = 128 channels of the M operand, N = BN channels of the N operand, K = 64 pixels / stage
// =====================================================================================================
constexpr int WG_PIX = 64;                      // pixels per stage (K extent)
constexpr int WG_BOX_BYTES = WG_PIX * 128;      // one (64 pixel x 64 channel) TMA box = 8 KB
constexpr int WG_THREADS = 192;

template <int BN, int MR = 128> struct WgCfg {
  static constexpr int M_BYTES = (MR / 64) * WG_BOX_BYTES;   // MR channels
  static constexpr int N_BYTES = (BN / 64) * WG_BOX_BYTES;
  static constexpr int STAGE = M_BYTES + N_BYTES;
  static constexpr int STAGES = BN >= 256 ? 4 : (BN >= 128 ? 5 : 6);
  static constexpr int SMEM = STAGES * STAGE + 1024 + 256;
};

// MR = 128: accumulator rows = 128 channels.  MR = 64: UMMA M = 64, whose 64 rows live in TMEM lanes
// {0-15, 32-47, 64-79, 96-111} (16 per warp quadrant).
template <int BN, int MR>
__global__ void __launch_bounds__(WG_THREADS)
    wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmN,
                    const __grid_constant__ WgradDev p) {
  using Cfg = WgCfg<BN, MR>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mblocks = p.cm / MR, nblocks = p.ngroup > 1 ? 1 : p.cn / BN;
  int wid = blockIdx.x;
  const int mb = wid % mblocks; wid /= mblocks;
  const int nb = wid % nblocks; wid /= nblocks;
  const int t = wid;                 // tap, or tap group when p.ngroup > 1
  const int split = blockIdx.y;
  const int tile_begin = split * p.tiles_per_split;
  const int tile_end = min(p.ntiles, tile_begin + p.tiles_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmM);
    tma_prefetch_desc(&tmN);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const bool grp = p.ngroup > 1;
      const int mdx = (p.m_tap && !grp) ? p.taps[t][0] : 0, mdy = (p.m_tap && !grp) ? p.taps[t][1] : 0,
                mdz = (p.m_tap && !grp) ? p.taps[t][2] : 0;
      const int ndx = (p.n_tap && !grp) ? p.taps[t][0] : 0, ndy = (p.n_tap && !grp) ? p.taps[t][1] : 0,
                ndz = (p.n_tap && !grp) ? p.taps[t][2] : 0;
      int stage = 0;
      uint32_t phase = 0;
      // (an L2 prefetch of the tiles ahead of the ring - cp.async.bulk.prefetch.tensor - was measured to slow every
      // geometry down by 1.4-1.8x: the extra TMA instructions compete with the loads themselves)
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        int x0, y0, n0;
        wg_tile_origin(p, tile, &x0, &y0, &n0);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (p.debug == 2 && tile >= tile_begin + Cfg::STAGES) {       // diagnostics: MMA pipeline without memory traffic
          mbar_arrive(&full_bar[stage]);
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
          continue;
        }
        mbar_expect_tx(&full_bar[stage], Cfg::STAGE);
        uint8_t* s = smem + stage * Cfg::STAGE;
#pragma unroll
        for (int j = 0; j < MR / 64; ++j)
          tma_load_5d(s + j * WG_BOX_BYTES, &tmM, &full_bar[stage], mb * MR + j * 64, x0 + mdx, y0 + mdy, mdz, n0);
        if (p.ngroup > 1) {
          const int bpt = (BN / 64) / p.ngroup;          // 64-channel boxes per tap
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) {
            const int tj = t * p.ngroup + j / bpt;
            // taps beyond the real count read far outside the tensor (zero fill): their accumulator columns are unused
            const bool real = tj < p.ntaps_total;
            const int sgn = p.grp_plus ? 1 : -1;
            const int sx = real ? x0 + sgn * p.taps[tj][0] : -(1 << 20);
            const int sy = real ? y0 + sgn * p.taps[tj][1] : 0;
            const int sz = (real && p.grp_plus) ? p.taps[tj][2] : 0;
            tma_load_5d(s + Cfg::M_BYTES + j * WG_BOX_BYTES, &tmN, &full_bar[stage], (j % bpt) * 64, sx, sy, sz, n0);
          }
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_5d(s + Cfg::M_BYTES + j * WG_BOX_BYTES, &tmN, &full_bar[stage], nb * BN + j * 64, x0 + ndx,
                        y0 + ndy, ndz, n0);
        }
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(MR, BN, 1, 1);   // both operands MN-major
      int stage = 0;
      uint32_t phase = 0;
      uint32_t first = 1;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t m_addr = smem_u32(smem + stage * Cfg::STAGE);
        const uint32_t n_addr = m_addr + Cfg::M_BYTES;
#pragma unroll
        for (int k = 0; k < WG_PIX / 16; ++k) {
          // MN-major SW128: 64-channel atoms LBO apart, 8-pixel groups SBO = 1024 B apart
          uint64_t da = umma_desc_sw128(m_addr + k * 2048, WG_BOX_BYTES, 1024);
          uint64_t db = umma_desc_sw128(n_addr + k * 2048, WG_BOX_BYTES, 1024);
          if (p.debug != 1) umma_bf16(tmem_base, da, db, idesc, first ? 0u : 1u);
          first = 0;
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tmem_full);
    }
  } else {
    const int q = warp & 3;
    const int m = MR == 128 ? mb * 128 + q * 32 + lane : mb * 64 + q * 16 + (lane & 15);
    const bool row_ok = MR == 128 || lane < 16;
    float* ws = p.ws + (((long long)split * p.ntaps + t) * p.cm + m) * p.cn + nb * BN;
    if (p.ngroup > 1 && tile_end > tile_begin) {
      // grouped taps: accumulator column c = (tap within the group, channel); cn channels per tap
      mbar_wait(tmem_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < BN; cc += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
        tmem_ld_wait();
        const int tj = t * p.ngroup + cc / p.cn, ch = cc % p.cn;
        if (row_ok && tj < p.ntaps_total) {
          float* w2 = p.ws + (((long long)split * p.ntaps_total + tj) * p.cm + m) * p.cn + ch;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(w2 + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                            __uint_as_float(v[j + 3]));
        }
      }
    } else if (p.ngroup > 1) {
      for (int cc = 0; cc < BN; cc += 32) {
        const int tj = t * p.ngroup + cc / p.cn, ch = cc % p.cn;
        if (row_ok && tj < p.ntaps_total) {
          float* w2 = p.ws + (((long long)split * p.ntaps_total + tj) * p.cm + m) * p.cn + ch;
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(w2 + j) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    } else if (tile_end > tile_begin) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < BN; cc += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(ws + cc + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                            __uint_as_float(v[j + 3]));
        }
      }
    } else if (row_ok) {
      for (int cc = 0; cc < BN; cc += 4) *reinterpret_cast<float4*>(ws + cc) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BN>(tmem_base);
  }
}

// dw[m*s_m + t*s_t + n*s_n] (+)= sum_s ws[s][t][m][n]   (optionally with one index split, see dwc_wgrad_t)
struct WgRemap {
  int axis, div, lo_limit, hi_limit;
  long long hi_stride, lo_stride;
};
// One 32 (m) x 32 (n) tile of one tap per block: the partial tiles are read along n (contiguous in the workspace) with
// four independent rows per thread in flight, summed over the splits in a fixed order, and written along whichever of
// m / n is the unit-stride axis of dw (through a shared-memory transpose when that is m).
__global__ void __launch_bounds__(256)
    wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int ntaps, int cm, int cn, float* dw, long long s_m,
                        long long s_t, long long s_n, int accumulate, WgRemap rm) {
  __shared__ float tile[32][33];
  const long long total = (long long)ntaps * cm * cn;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int t = blockIdx.z, m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  {
    const int n = n0 + tx;
    const float* q[4];
    bool ok[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = m0 + ty + 8 * r;
      ok[r] = m < cm && n < cn;
      q[r] = ws + ((long long)t * cm + (ok[r] ? m : 0)) * cn + (ok[r] ? n : 0);
    }
    int k = 0;
    for (; k + 4 <= splits; k += 4) {
      float v[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) v[j][r] = ok[r] ? __ldg(q[r] + (k + j) * total) : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) s[r] += v[j][r];
    }
    for (; k < splits; ++k)
#pragma unroll
      for (int r = 0; r < 4; ++r) s[r] += ok[r] ? __ldg(q[r] + k * total) : 0.f;
  }
  const bool along_m = s_m < s_n;            // which axis is contiguous in dw
  if (along_m) {
#pragma unroll
    for (int r = 0; r < 4; ++r) tile[ty + 8 * r][tx] = s[r];
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int m = along_m ? m0 + tx : m0 + ty + 8 * r;
    const int n = along_m ? n0 + ty + 8 * r : n0 + tx;
    if (m >= cm || n >= cn) continue;
    const float v = along_m ? tile[tx][ty + 8 * r] : s[r];
    long long om = m * s_m, on = n * s_n;
    if (rm.axis == 1) {
      if (m % rm.div >= rm.lo_limit || m / rm.div >= rm.hi_limit) continue;
      om = (m / rm.div) * rm.hi_stride + (m % rm.div) * rm.lo_stride;
    } else if (rm.axis == 2) {
      if (n % rm.div >= rm.lo_limit || n / rm.div >= rm.hi_limit) continue;
      on = (n / rm.div) * rm.hi_stride + (n % rm.div) * rm.lo_stride;
    }
    float* o = dw + om + t * s_t + on;
    *o = accumulate ? (*o + v) : v;
  }
}

// bias gradient: column sums of dY over all pixels, two deterministic stages.
// grid (ceil(C/32), DB_SPLITS); block 256 = 4 channel-vectors (8 channels each) x 64 pixel lanes; 16-byte loads.
constexpr int DB_SPLITS = 256;        // maximum; the launch picks enough splits to fill the GPU
template <typename T>
__global__ void __launch_bounds__(256) dbias_partial_kernel(const T* __restrict__ A, long long sx, long long sy, long long sn,
                                                            int W, int H, int N, float* part, int ca, int nsplit) {
  __shared__ float red[64][33];
  const int cv = threadIdx.x & 3, pl = threadIdx.x >> 2;
  const int c0 = blockIdx.x * 32 + cv * 8;
  const long long total = (long long)N * H * W;
  const long long per = (total + nsplit - 1) / nsplit;
  const long long p0 = blockIdx.y * per, p1 = min(total, p0 + per);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (c0 < ca) {
    for (long long p = p0 + pl; p < p1; p += 64) {
      int x = (int)(p % W);
      long long r = p / W;
      int y = (int)(r % H);
      int n = (int)(r / H);
      const T* q = A + n * sn + y * sy + x * sx + c0;
      if (c0 + 8 <= ca) {
        float v[8];
        Vec8<T>::load(q, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += v[e];
      } else {
        for (int e = 0; e < 8 && c0 + e < ca; ++e) acc[e] += to_f<T>(q[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[pl][cv * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = 0.f;
    for (int i = 0; i < 64; ++i) t += red[i][threadIdx.x];
    int c = blockIdx.x * 32 + threadIdx.x;
    if (c < ca) part[(long long)blockIdx.y * ca + c] = t;
  }
}
__global__ void dbias_final_kernel(const float* part, int ca, float* dbias, int accumulate, int nsplit) {
  int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= ca) return;
  float s = 0.f;
  for (int k = 0; k < nsplit; ++k) s += part[(long long)k * ca + ch];
  dbias[ch] = accumulate ? dbias[ch] + s : s;
}

// =====================================================================================================
// host
// =====================================================================================================
struct WgPlan {
  bool swap;
  int cm, cn, bn, mr, splits, tiles_per_split, ntiles, items;
  int ngroup;                       // > 1: tap-grouped N operand (see WgradDev::ngroup)
  int grp_plus;                     // see WgradDev::grp_plus
  int box[3], tiles[3];             // pixel boxes / tile counts actually used (grouped mode re-tiles the M grid)
};

static int wg_plan(const dwc_wgrad_t* g, WgPlan* pl) {
  pl->ntiles = g->tiles[0] * g->tiles[1] * g->tiles[2];
  pl->ngroup = 1;
  for (int i = 0; i < 3; ++i) { pl->box[i] = g->box[i]; pl->tiles[i] = g->tiles[i]; }
  if (g->backend == DWC_TC) {
    // accumulator rows need a multiple of 128 channels
    DWC_CHECK(g->ca % 64 == 0 && g->cb % 64 == 0, "dwc_wgrad: tcgen05 needs channel counts (%d,%d) %% 64 == 0", g->ca,
              g->cb);
    // accumulator rows (M) come from one operand's channels, columns (N) from the other's: rows need a multiple of 128,
    // and the widest tile (N = 256) has the best operand reuse - swap when that puts 256 channels on N
    pl->swap = ((g->ca % 128 != 0) && (g->cb % 128 == 0)) ||
               (g->ca % 256 == 0 && g->cb % 128 == 0 && g->cb % 256 != 0 && g->remap_axis == 0);
    pl->cm = pl->swap ? g->cb : g->ca;
    pl->cn = pl->swap ? g->ca : g->cb;
    pl->mr = pl->cm % 128 == 0 ? 128 : 64;
    pl->bn = pl->cn % 256 == 0 ? 256 : (pl->cn % 128 == 0 ? 128 : 64);
    if (pl->mr == 64) pl->bn = 64;
    pl->items = (pl->cm / pl->mr) * (pl->cn / pl->bn) * g->ntaps;
    // Few output channels (cout 64 / 128 after the operand swap: N would be 64 / 128 wide and every M = 128 MMA is
    // bound by its shared-memory A read): put 256 / cout taps side by side in N.  dW[co,tap,ci] = sum_p' dY[p'-tap,co]
    // X[p',ci]: X is read untapped over ITS (padded) pixel grid, dY at p' - tap (zero outside, by TMA fill).
    static int grp_on = -1;
    if (grp_on < 0) {
      const char* e = getenv("DWC_WGRAD_GROUP");
      grp_on = e ? atoi(e) : 1;
    }
    bool plain = g->remap_axis == 0 && g->ntaps >= 4;
    for (int t = 0; t < g->ntaps && plain; ++t)
      plain = g->taps[t * 3 + 2] == 0 && g->taps[t * 3] >= 0 && g->taps[t * 3 + 1] >= 0;
    pl->grp_plus = 0;
    if (grp_on && !pl->swap && g->remap_axis == 0 && pl->mr == 128 && (pl->cn == 64 || pl->cn == 128) &&
        g->ntaps >= 256 / pl->cn) {
      // N operand = the tapped input with few channels (64 -> 128, 64 -> 64 layers): 256 / cn taps side by side in N,
      // each box read at its own tap offset / parity plane; M = dY untapped, tiles over dY's grid as in the classic mode
      pl->ngroup = 256 / pl->cn;
      pl->bn = 256;
      pl->grp_plus = 1;
      pl->items = (pl->cm / 128) * cdiv(g->ntaps, pl->ngroup);
    } else
    if (grp_on && pl->swap && plain && pl->mr == 128 && (pl->cn == 64 || pl->cn == 128) && g->b_dim[3] == 1) {
      pl->ngroup = 256 / pl->cn;
      pl->bn = 256;
      pl->items = (pl->cm / 128) * cdiv(g->ntaps, pl->ngroup);
      // 64-pixel boxes over the M operand's grid (b = padded input) with the least padding
      const long long W = g->b_dim[1], H = g->b_dim[2];
      long long best = -1;
      for (int bx = 64; bx >= 8; bx >>= 1) {
        const int by = 64 / bx;
        const long long area = (long long)cdiv(W, bx) * bx * cdiv(H, by) * by;
        if (best < 0 || area < best) {
          best = area;
          pl->box[0] = bx; pl->box[1] = by; pl->box[2] = 1;
        }
      }
      pl->tiles[0] = cdiv(W, pl->box[0]); pl->tiles[1] = cdiv(H, pl->box[1]); pl->tiles[2] = (int)g->b_dim[4];
      pl->ntiles = pl->tiles[0] * pl->tiles[1] * pl->tiles[2];
    }
  } else {
    pl->swap = false;
    pl->cm = g->ca;
    pl->cn = g->cb;
    pl->bn = 64;
    pl->mr = 64;
    pl->items = cdiv(pl->cm, 64) * cdiv(pl->cn, 64) * g->ntaps;
  }
  // one wave of CTAs: fewer K splits mean fewer partial tiles for the (HBM-bound) reduction pass
  int want = dwc_num_sms() / pl->items;
  if (want < 1) want = 1;
  if (want > pl->ntiles) want = pl->ntiles;
  if (want > 64) want = 64;
  pl->tiles_per_split = cdiv(pl->ntiles, want);
  pl->splits = cdiv(pl->ntiles, pl->tiles_per_split);
  return 0;
}

extern "C" int64_t dwc_wgrad_workspace_bytes(const dwc_wgrad_t* g) {
  WgPlan pl;
  if (wg_plan(g, &pl)) return -1;
  int64_t ws = (int64_t)pl.splits * g->ntaps * pl.cm * pl.cn * 4;
  ws += (int64_t)DB_SPLITS * g->ca * 4;
  return ws;
}

template <int BN, int MR>
static int launch_wg_tc(const dwc_wgrad_t* g, const WgradDev& d, const WgPlan& pl, cudaStream_t st) {
  using Cfg = WgCfg<BN, MR>;
  CUtensorMap tmM, tmN;
  const bool sw = pl.swap;
  if (dwc_make_tmap5(&tmM, sw ? g->b : g->a, sw ? g->b_dim : g->a_dim, sw ? g->b_str : g->a_str, pl.box[0], pl.box[1],
                     1, pl.box[2]))
    return 1;
  if (dwc_make_tmap5(&tmN, sw ? g->a : g->b, sw ? g->a_dim : g->b_dim, sw ? g->a_str : g->b_str, pl.box[0], pl.box[1],
                     1, pl.box[2]))
    return 1;
  static bool attr_set = false;
  if (!attr_set) {
    DWC_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<BN, MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  dim3 grid(pl.items, pl.splits);
  wgrad_tc_kernel<BN, MR><<<grid, WG_THREADS, Cfg::SMEM, st>>>(tmM, tmN, d);
  DWC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dwc_wgrad(const dwc_wgrad_t* g, dwc_stream_t stream) {
  DWC_CHECK(g != nullptr, "dwc_wgrad: null params");
  DWC_CHECK(g->ntaps > 0 && g->ntaps <= DWC_MAX_TAPS, "dwc_wgrad: ntaps out of range");
  WgPlan pl;
  if (wg_plan(g, &pl)) return 1;
  const int64_t need = dwc_wgrad_workspace_bytes(g);
  DWC_CHECK(g->workspace != nullptr && g->workspace_bytes >= need, "dwc_wgrad: workspace too small (%lld < %lld)",
            (long long)g->workspace_bytes, (long long)need);
  if (g->backend == DWC_TC)
    DWC_CHECK(g->box[0] * g->box[1] * g->box[2] == WG_PIX && g->dtype == DWC_BF16,
              "dwc_wgrad: tcgen05 needs bf16 and 64-pixel boxes");
  WgradDev d;
  memset(&d, 0, sizeof(d));
  const bool sw = pl.swap;
  d.m_ptr = sw ? g->b : g->a;
  d.n_ptr = sw ? g->a : g->b;
  for (int i = 0; i < 5; ++i) {
    d.m_dim[i] = sw ? g->b_dim[i] : g->a_dim[i];
    d.m_str[i] = sw ? g->b_str[i] : g->a_str[i];
    d.n_dim[i] = sw ? g->a_dim[i] : g->b_dim[i];
    d.n_str[i] = sw ? g->a_str[i] : g->b_str[i];
  }
  d.m_tap = sw ? 1 : 0;
  d.n_tap = sw ? 0 : 1;
  d.cm = pl.cm; d.cn = pl.cn;
  d.box_x = pl.box[0]; d.box_y = pl.box[1]; d.box_n = pl.box[2];
  d.tiles_x = pl.tiles[0]; d.tiles_y = pl.tiles[1]; d.tiles_n = pl.tiles[2];
  d.ntiles = pl.ntiles; d.splits = pl.splits; d.tiles_per_split = pl.tiles_per_split;
  d.ntaps = pl.ngroup > 1 ? cdiv(g->ntaps, pl.ngroup) : g->ntaps;
  d.ngroup = pl.ngroup;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("DWC_WGRAD_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    d.debug = dbg;
  }
  d.grp_plus = pl.grp_plus;
  d.ntaps_total = g->ntaps;
  d.ws = g->workspace;
  for (int t = 0; t < g->ntaps; ++t)
    for (int j = 0; j < 3; ++j) d.taps[t][j] = g->taps[t * 3 + j];
  cudaStream_t st = as_stream(stream);

  if (g->backend == DWC_TC) {
    int rc;
    if (pl.mr == 64) rc = launch_wg_tc<64, 64>(g, d, pl, st);
    else if (pl.bn == 256) rc = launch_wg_tc<256, 128>(g, d, pl, st);
    else if (pl.bn == 128) rc = launch_wg_tc<128, 128>(g, d, pl, st);
    else rc = launch_wg_tc<64, 128>(g, d, pl, st);
    if (rc) return rc;
  } else {
    dim3 grid(pl.items, pl.splits);
    if (g->dtype == DWC_F32) wgrad_simt_kernel<float><<<grid, 256, 0, st>>>(d);
    else wgrad_simt_kernel<bf16><<<grid, 256, 0, st>>>(d);
    DWC_LAUNCH_CHECK();
  }
  WgRemap rm;
  rm.axis = g->remap_axis == 0 ? 0 : ((g->remap_axis == 1) != sw ? 1 : 2);   // axis in (m, n) terms after the operand swap
  rm.div = g->remap_div > 0 ? g->remap_div : 1;
  rm.lo_limit = g->remap_lo_limit; rm.hi_limit = g->remap_hi_limit;
  rm.hi_stride = g->remap_hi_stride; rm.lo_stride = g->remap_lo_stride;
  const long long total = (long long)g->ntaps * pl.cm * pl.cn;
  DWC_CHECK(g->ntaps <= 65535 && cdiv(pl.cm, 32) <= 65535, "dwc_wgrad: reduce grid too large");
  wgrad_reduce_kernel<<<dim3(cdiv(pl.cn, 32), cdiv(pl.cm, 32), g->ntaps), 256, 0, st>>>(
      g->workspace, pl.splits, g->ntaps, pl.cm, pl.cn, g->dw, sw ? g->s_b : g->s_a, g->s_t, sw ? g->s_a : g->s_b,
      g->accumulate, rm);
  DWC_LAUNCH_CHECK();
  if (g->dbias) {
    float* part = g->workspace + (int64_t)pl.splits * g->ntaps * pl.cm * pl.cn;
    const long long npix = (long long)g->a_dim[1] * g->a_dim[2] * g->a_dim[4];
    int nsplit = cdiv(4 * dwc_num_sms(), cdiv(g->ca, 32));
    if (nsplit > DB_SPLITS) nsplit = DB_SPLITS;
    if (nsplit > cdiv(npix, 64)) nsplit = cdiv(npix, 64);
    if (nsplit < 1) nsplit = 1;
    dim3 grid(cdiv(g->ca, 32), nsplit);
    DWC_CHECK(g->ca % 8 == 0 || g->ca < 8, "dwc_wgrad: dbias needs ca %% 8 == 0");
    if (g->dtype == DWC_F32)
      dbias_partial_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(g->a), g->a_str[1], g->a_str[2],
                                                        g->a_str[4], (int)g->a_dim[1], (int)g->a_dim[2],
                                                        (int)g->a_dim[4], part, g->ca, nsplit);
    else
      dbias_partial_kernel<bf16><<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(g->a), g->a_str[1], g->a_str[2],
                                                       g->a_str[4], (int)g->a_dim[1], (int)g->a_dim[2],
                                                       (int)g->a_dim[4], part, g->ca, nsplit);
    DWC_LAUNCH_CHECK();
    dbias_final_kernel<<<cdiv(g->ca, 128), 128, 0, st>>>(part, g->ca, g->dbias, g->accumulate, nsplit);
    DWC_LAUNCH_CHECK();
  }
  return 0;
}
