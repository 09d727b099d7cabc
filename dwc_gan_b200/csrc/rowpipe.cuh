// Row-streaming bf16 versions of the HBM-bound norm passes (statistics, normalise + activation + residual + reflect pad,
// backward reductions, backward apply).
//
// The per-thread 16-byte-load kernels in elementwise.cu are latency bound (ncu: 10-35 % of DRAM throughput at 20-40 %
// occupancy, ~128 registers): the bytes a thread can keep in flight are tied to its registers.  Here the loads are 1-D
// bulk copies (cp.async.bulk, the TMA engine without a tensor map) of whole image rows - W*C contiguous bf16, 16 KB at
// every resolution of the 128x128 network - into a shared-memory ring guarded by mbarriers, issued by one thread
// several rows ahead.  The 256 threads of the CTA consume a row from shared memory (conflict-free 16-byte reads), keep
// their per-channel coefficients in registers and write results with fully contiguous 16-byte stores.
// grid (row splits, N); a CTA owns a contiguous range of rows of one sample, all channels.
#pragma once
#include "common.cuh"

enum { RM_STATS = 0, RM_BRED = 1, RM_FWD = 2, RM_BAPPLY = 3 };

struct RowP {
  HB y;            // primary input (plain layout, any halo): conv output
  HB d;            // RM_BRED / RM_BAPPLY: upstream gradient (reflections already folded, or no halo); RM_FWD: residual
  HB o1;           // RM_FWD: output (reflect halo, plain or parity planes); RM_BAPPLY: dy (zero halo)
  HB o2;           // RM_BAPPLY: dres (zero halo)
  const float4* coef;   // per (n,c): scale, shift, (mean, rstd)
  const float4* bco;    // per (n,c): a, b, c of dy = a*dz + b*y + c
  float2* part;         // RM_STATS / RM_BRED: [n][split][c]
  int act, has_d, has_o2;
  int nseg, segw, segbytes, stages;
  // optional in-kernel coefficient computation (replaces the separate finalize launches): kind 1 IN, 2 AdaIN, 3 LN
  const float2* nstats;    // RM_FWD: per-(n, split, c) {sum, sum of squares}; RM_BAPPLY: {sum dz, sum dz*y}; null = use coef/bco
  const float* nweight;    // AdaIN [N,C] / LN [C] scale
  const float* nbias;      // AdaIN [N,C] / LN [C] shift (RM_FWD)
  float4* coef_out;        // RM_FWD: {scale, shift, mean, rstd} for the backward pass
  float* dweight;          // RM_BAPPLY, AdaIN: [N,C] parameter gradients (overwritten)
  float* dbias;
  int nkind, nsplits, nhw;
  float neps;
  int dbytes;      // bytes of the second operand per stage (0: none).  d in parity-plane layout: the two plane rows of the
  int d_planes;    // padded row (even X, odd X), wq*C elements each, loaded whole (nseg == 1)
  int d_fold;      // RM_BRED: d still carries the gradient of its reflect halo; the kernel loads whole PADDED rows, folds
                   // the mirror images into the band pixels on the fly (same order and bf16 rounding as dwc_fold_halo)
                   // and writes the folded band back, so that the apply pass reads a folded buffer (nseg == 1, plain d)
};

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void rp_unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 rp_pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
__device__ __forceinline__ float rp_act_fwd(float z, int act) {
  if (act == 1) return z > 0.f ? z : 0.f;
  if (act == 2) return z > 0.f ? z : 0.1f * z;
  return z;
}
__device__ __forceinline__ float rp_act_grad(float z, int act) {
  if (act == 1) return z > 0.f ? 1.f : 0.f;
  if (act == 2) return z > 0.f ? 1.f : 0.1f;
  return 1.f;
}

// zero the halo of row `iy` (left / right pixels), and the top / bottom halo rows when iy is the first / last row
__device__ __forceinline__ void rp_zero_halo(const HB& b, int n, int iy, int cvs, int tid) {
  if (b.halo == 0) return;
  bf16* base = reinterpret_cast<bf16*>(b.ptr);
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  const int side = b.halo * cvs;
  for (int q = tid; q < 2 * side; q += 256) {
    const int r = q >= side ? q - side : q;
    const int X = (q >= side ? b.halo + b.w : 0) + r / cvs;
    *reinterpret_cast<uint4*>(base + b.off_padded(n, iy + b.halo, X) + (r % cvs) * 8) = z;
  }
  const int rowchunks = b.wp * cvs;
  if (iy == 0) {
    for (int q = tid; q < b.halo * rowchunks; q += 256) {
      const int Y = q / rowchunks, r = q - Y * rowchunks;
      *reinterpret_cast<uint4*>(base + b.off_padded(n, Y, r / cvs) + (r % cvs) * 8) = z;
    }
  }
  if (iy == b.h - 1) {
    for (int q = tid; q < b.halo * rowchunks; q += 256) {
      const int Y = q / rowchunks, r = q - Y * rowchunks;
      *reinterpret_cast<uint4*>(base + b.off_padded(n, b.halo + b.h + Y, r / cvs) + (r % cvs) * 8) = z;
    }
  }
}

// Forward coefficients of sample n for all C channels -> shared memory (and global memory for the backward pass when
// `write`): the arithmetic of norm_finalize_kernel, executed by every CTA of the sample on the (L2-resident) partial sums.
__device__ __forceinline__ void rp_fwd_coef(const RowP& p, int n, int C, float2* sco, double* dsm, bool write) {
  const int tid = threadIdx.x, splits = p.nsplits, hw = p.nhw;
  if (p.nkind == 3) {
    double s = 0, q = 0;
    for (int c = tid; c < C; c += 256)
      for (int k = 0; k < splits; ++k) {
        const float2 v = p.nstats[((long long)n * splits + k) * C + c];
        s += v.x; q += v.y;
      }
    s = block_sum_d(s, dsm);
    q = block_sum_d(q, dsm);
    const double M = (double)C * hw;
    const double mean = s / M;
    double var = (q - M * mean * mean) / (M - 1.0);
    if (var < 0) var = 0;
    const float inv = (float)(1.0 / (sqrt(var) + (double)p.neps));
    for (int c = tid; c < C; c += 256) {
      const float g = p.nweight[c], b = p.nbias[c];
      const float4 cf = make_float4(g * inv, b - (float)mean * g * inv, (float)mean, inv);
      sco[c] = make_float2(cf.x, cf.y);
      if (write) p.coef_out[(long long)n * C + c] = cf;
    }
  } else {
    for (int c = tid; c < C; c += 256) {
      double s = 0, q = 0;
      for (int k = 0; k < splits; ++k) {
        const float2 v = p.nstats[((long long)n * splits + k) * C + c];
        s += v.x; q += v.y;
      }
      const double mean = s / hw;
      double var = q / hw - mean * mean;
      if (var < 0) var = 0;
      const float rstd = (float)(1.0 / sqrt(var + (double)p.neps));
      const float w = p.nkind == 2 ? p.nweight[(long long)n * C + c] : 1.f;
      const float b = p.nkind == 2 ? p.nbias[(long long)n * C + c] : 0.f;
      const float4 cf = make_float4(w * rstd, b - (float)mean * rstd * w, (float)mean, rstd);
      sco[c] = make_float2(cf.x, cf.y);
      if (write) p.coef_out[(long long)n * C + c] = cf;
    }
  }
  __syncthreads();
}

// Backward coefficients {a, b, c} of dy = a*dz + b*y + c (norm_bwd_finalize_kernel's arithmetic) -> shared memory;
// AdaIN parameter gradients are written by the sample's first CTA.
__device__ __forceinline__ void rp_bwd_coef(const RowP& p, int n, int C, float4* sbco, double* dsm, bool write) {
  const int tid = threadIdx.x, splits = p.nsplits, hw = p.nhw;
  if (p.nkind == 3) {
    const double M = (double)C * hw;
    const double mean = p.coef[(long long)n * C].z;
    double g1 = 0, g2 = 0;
    for (int c = tid; c < C; c += 256) {
      double S1 = 0, S2 = 0;
      for (int k = 0; k < splits; ++k) {
        const float2 v = p.nstats[((long long)n * splits + k) * C + c];
        S1 += v.x; S2 += v.y;
      }
      const double g = p.nweight[c];
      g1 += g * S1;
      g2 += g * (S2 - mean * S1);
    }
    g1 = block_sum_d(g1, dsm);
    g2 = block_sum_d(g2, dsm);
    const double inv = p.coef[(long long)n * C].w;
    const double sd = 1.0 / inv - (double)p.neps;
    const double K = sd > 0 ? g2 * inv * inv / ((M - 1.0) * sd) : 0.0;
    const double b = -K;
    const double cc = -g1 * inv / M + K * mean;
    for (int c = tid; c < C; c += 256) sbco[c] = make_float4((float)(p.nweight[c] * inv), (float)b, (float)cc, 0.f);
  } else {
    for (int c = tid; c < C; c += 256) {
      double S1 = 0, S2 = 0;
      for (int k = 0; k < splits; ++k) {
        const float2 v = p.nstats[((long long)n * splits + k) * C + c];
        S1 += v.x; S2 += v.y;
      }
      const float4 q = p.coef[(long long)n * C + c];
      const double mean = q.z, rstd = q.w;
      const double w = p.nkind == 2 ? (double)p.nweight[(long long)n * C + c] : 1.0;
      const double m1 = S1 / hw;
      const double m2 = rstd * (S2 / hw - mean * m1);
      if (p.nkind == 2 && write) {
        p.dbias[(long long)n * C + c] = (float)S1;
        p.dweight[(long long)n * C + c] = (float)(rstd * (S2 - mean * S1));
      }
      const double a = w * rstd;
      const double b = -rstd * rstd * w * m2;
      const double cc = -rstd * w * m1 - b * mean;
      sbco[c] = make_float4((float)a, (float)b, (float)cc, 0.f);
    }
  }
  __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(256) row_kernel(const __grid_constant__ RowP p) {
  extern __shared__ __align__(128) uint8_t rsm[];
  __shared__ uint64_t full[8];
  const int tid = threadIdx.x;
  const int n = blockIdx.y;
  const int H = p.y.h, W = p.y.w, C = p.y.c;
  const int cvs = C >> 3, cv = tid % cvs, c0 = cv * 8;
  const int nb = MODE == RM_STATS ? 1 : (MODE == RM_FWD ? (p.has_d ? 2 : 1) : 2);
  const int S = p.stages;
  // balanced contiguous row ranges (they differ by at most one row)
  const int r0 = (int)(((long long)blockIdx.x * H) / gridDim.x), r1 = (int)(((long long)(blockIdx.x + 1) * H) / gridDim.x);
  const int cnt = (r1 - r0) * p.nseg;
  const int stage_bytes = p.segbytes + (nb == 2 ? p.dbytes : 0);
  const bf16* yb = reinterpret_cast<const bf16*>(p.y.ptr);
  const bf16* db = reinterpret_cast<const bf16*>(p.d.ptr);

  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  pdl_prologue();
  __syncthreads();
  auto issue = [&](int i) {
    const int u = r0 * p.nseg + i;
    const int row = u / p.nseg, x0 = (u - row * p.nseg) * p.segw;
    const int s = i % S;
    uint8_t* dst = rsm + (size_t)s * stage_bytes;
    mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
    bulk_load_1d(dst, yb + p.y.off(n, row, x0), (uint32_t)p.segbytes, &full[s]);
    if (nb == 2) {
      if (p.d_planes) {
        const int Y = row + p.d.halo;
        bulk_load_1d(dst + p.segbytes, db + p.d.off_padded(n, Y, 0), (uint32_t)(p.dbytes >> 1), &full[s]);
        bulk_load_1d(dst + p.segbytes + (p.dbytes >> 1), db + p.d.off_padded(n, Y, 1), (uint32_t)(p.dbytes >> 1),
                     &full[s]);
      } else if (MODE == RM_BRED && p.d_fold) {
        bulk_load_1d(dst + p.segbytes, db + p.d.off_padded(n, row + p.d.halo, 0), (uint32_t)p.dbytes, &full[s]);
      } else {
        bulk_load_1d(dst + p.segbytes, db + p.d.off(n, row, x0), (uint32_t)p.segbytes, &full[s]);
      }
    }
  };
  // shared-memory byte offset (inside the d part of a stage) of chunk q = (pixel q / cvs, channel group cv)
  auto d_off = [&](int q) -> size_t {
    if (!p.d_planes) return (size_t)q * 16;
    const int X = q / cvs + p.d.halo;
    return ((size_t)((X & 1) * (p.d.wp >> 1) + (X >> 1)) * C + c0) * 2;
  };
  if (tid == 0)
    for (int i = 0; i < S - 1 && i < cnt; ++i) issue(i);

  // per-channel coefficients of this thread's 8 channels
  float sc[8], sh[8], ba[8], bb[8], bc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc[e] = 1.f; sh[e] = 0.f; ba[e] = 1.f; bb[e] = 0.f; bc[e] = 0.f;
  }
  __shared__ float4 s_co[512];                          // in-kernel coefficients (C <= 512)
  __shared__ double s_d[256];
  if (MODE == RM_FWD && p.nstats) {
    rp_fwd_coef(p, n, C, reinterpret_cast<float2*>(s_co), s_d, blockIdx.x == 0);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float2 q = reinterpret_cast<const float2*>(s_co)[c0 + e];
      sc[e] = q.x; sh[e] = q.y;
    }
  } else if (MODE != RM_STATS && p.coef) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float4 q = __ldg(p.coef + (long long)n * C + c0 + e);
      sc[e] = q.x; sh[e] = q.y;
    }
  }
  if (MODE == RM_BAPPLY && p.nstats) {
    rp_bwd_coef(p, n, C, s_co, s_d, blockIdx.x == 0);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float4 q = s_co[c0 + e];
      ba[e] = q.x; bb[e] = q.y; bc[e] = q.z;
    }
  } else if (MODE == RM_BAPPLY && p.bco) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float4 q = __ldg(p.bco + (long long)n * C + c0 + e);
      ba[e] = q.x; bb[e] = q.y; bc[e] = q.z;
    }
  }
  float a0[8], a1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a0[e] = a1[e] = 0.f;

  const int nchunks = p.segbytes >> 4;
  const int act = p.act;
  for (int i = 0; i < cnt; ++i) {
    __syncthreads();                                   // everyone is done with unit i-1: its stage may be refilled
    if (tid == 0 && i + S - 1 < cnt) issue(i + S - 1);
    const int s = i % S;
    mbar_wait(&full[s], (uint32_t)((i / S) & 1));
    const uint8_t* sy = rsm + (size_t)s * stage_bytes;
    const uint8_t* sd = sy + p.segbytes;
    const int u = r0 * p.nseg + i;
    const int row = u / p.nseg, x0 = (u - row * p.nseg) * p.segw;
    if (MODE == RM_STATS) {
#pragma unroll 4
      for (int q = tid; q < nchunks; q += 256) {
        float v[8];
        rp_unpack8(*reinterpret_cast<const uint4*>(sy + (size_t)q * 16), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) { a0[e] += v[e]; a1[e] += v[e] * v[e]; }
      }
    } else if (MODE == RM_BRED && p.d_fold) {
      // d row = whole padded row [X = 0 .. W + 2h); interior pixel px sits at X = px + h.
      // Phase 1: the band pixels (h columns next to each edge; the whole row when it is a band row) take the sum of
      // their mirror images, in dwc_fold_halo's order and rounding, in the staged row AND in global memory.
      // Phase 2: the plain reduction over the (now folded) staged row.
      const int h = p.d.halo;
      const int Ym = (row >= 1 && row <= h) ? h - row : ((row >= H - 1 - h && row <= H - 2) ? h + 2 * (H - 1) - row : -1);
      bf16* dg = reinterpret_cast<bf16*>(p.d.ptr);
      const long long rowY = p.d.off_padded(n, row + h, 0);
      const long long rowM = Ym >= 0 ? p.d.off_padded(n, Ym, 0) : 0;
      uint8_t* sdw = const_cast<uint8_t*>(sd);
      const int nband = Ym >= 0 ? W * cvs : 2 * h * cvs;
      for (int j = tid; j < nband; j += 256) {
        const int bp = j / cvs, bc = (j - bp * cvs) * 8;
        const int px = Ym >= 0 ? bp : (bp < h ? 1 + bp : W - 1 - h + (bp - h));
        const int xm = (px >= 1 && px <= h) ? h - px : ((px >= W - 1 - h && px <= W - 2) ? h + 2 * (W - 1) - px : -1);
        float g[8], t[8];
        uint8_t* slot = sdw + ((size_t)(px + h) * C + bc) * 2;
        rp_unpack8(*reinterpret_cast<const uint4*>(slot), g);
        if (xm >= 0) {
          rp_unpack8(*reinterpret_cast<const uint4*>(sd + ((size_t)xm * C + bc) * 2), t);
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] += t[e];
        }
        if (Ym >= 0) {
          rp_unpack8(__ldg(reinterpret_cast<const uint4*>(dg + rowM + (long long)(px + h) * C + bc)), t);
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] += t[e];
          if (xm >= 0) {
            rp_unpack8(__ldg(reinterpret_cast<const uint4*>(dg + rowM + (long long)xm * C + bc)), t);
#pragma unroll
            for (int e = 0; e < 8; ++e) g[e] += t[e];
          }
        }
        const uint4 folded = rp_pack8(g);
        *reinterpret_cast<uint4*>(slot) = folded;
        *reinterpret_cast<uint4*>(dg + rowY + (long long)(px + h) * C + bc) = folded;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes to a stage the bulk copies refill later
      __syncthreads();
      const uint8_t* sdi = sd + (size_t)h * C * 2;       // interior part of the padded row
#pragma unroll 4
      for (int q = tid; q < nchunks; q += 256) {
        float v[8], g[8];
        rp_unpack8(*reinterpret_cast<const uint4*>(sy + (size_t)q * 16), v);
        rp_unpack8(*reinterpret_cast<const uint4*>(sdi + (size_t)q * 16), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float dz = g[e] * rp_act_grad(sc[e] * v[e] + sh[e], act);
          a0[e] += dz; a1[e] += dz * v[e];
        }
      }
    } else if (MODE == RM_BRED) {
#pragma unroll 4
      for (int q = tid; q < nchunks; q += 256) {
        float v[8], g[8];
        rp_unpack8(*reinterpret_cast<const uint4*>(sy + (size_t)q * 16), v);
        rp_unpack8(*reinterpret_cast<const uint4*>(sd + d_off(q)), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float dz = g[e] * rp_act_grad(sc[e] * v[e] + sh[e], act);
          a0[e] += dz; a1[e] += dz * v[e];
        }
      }
    } else if (MODE == RM_FWD) {
      bf16* ob = reinterpret_cast<bf16*>(p.o1.ptr);
      const int halo = p.o1.halo;
      // destination rows: the row itself and up to two reflections of it in the halo
      const bool r1 = halo > 0 && row >= 1 && row <= halo, r2 = halo > 0 && row >= H - 1 - halo && row <= H - 2;
      const int Ya = row + halo;
      const int Yb = r1 ? halo - row : (r2 ? halo + 2 * (H - 1) - row : -1);
      const int Yc = (r1 && r2) ? halo + 2 * (H - 1) - row : -1;
      const int ny = 1 + (Yb >= 0) + (Yc >= 0);
      // Output addressing hoisted out of the chunk loop: per destination row Y the element offset of padded column
      // X is ybase[Y][X & pmask] + (X >> pshift) * C (plain layout: pmask = pshift = 0; parity planes: 1 / 1)
      const int pmask = p.o1.layout ? 1 : 0;
      // (explicit scalars instead of small arrays: dynamically indexed arrays would live in local memory)
      const long long ya0 = p.o1.off_padded(n, Ya, 0) + c0;
      const long long ya1 = p.o1.layout ? p.o1.off_padded(n, Ya, 1) + c0 : ya0;
      const long long yb0 = ny > 1 ? p.o1.off_padded(n, Yb, 0) + c0 : 0;
      const long long yb1 = ny > 1 ? (p.o1.layout ? p.o1.off_padded(n, Yb, 1) + c0 : yb0) : 0;
      const long long yc0 = ny > 2 ? p.o1.off_padded(n, Yc, 0) + c0 : 0;
      const long long yc1 = ny > 2 ? (p.o1.layout ? p.o1.off_padded(n, Yc, 1) + c0 : yc0) : 0;
      auto put = [&](int X, const uint4& o) {
        const long long xo = (long long)(X >> pmask) * C;
        const bool odd = (X & pmask) != 0;
        *reinterpret_cast<uint4*>(ob + (odd ? ya1 : ya0) + xo) = o;
        if (ny > 1) *reinterpret_cast<uint4*>(ob + (odd ? yb1 : yb0) + xo) = o;
        if (ny > 2) *reinterpret_cast<uint4*>(ob + (odd ? yc1 : yc0) + xo) = o;
      };
      const int pstep = 256 / cvs;                     // pixels between two chunks of this thread (cvs divides 256)
      int px = x0 + tid / cvs;
#pragma unroll 4
      for (int q = tid; q < nchunks; q += 256, px += pstep) {
        float v[8];
        rp_unpack8(*reinterpret_cast<const uint4*>(sy + (size_t)q * 16), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = rp_act_fwd(sc[e] * v[e] + sh[e], act);
        if (p.has_d) {
          float r[8];
          rp_unpack8(*reinterpret_cast<const uint4*>(sd + (size_t)q * 16), r);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] += r[e];
        }
        const uint4 o = rp_pack8(v);
        put(px + halo, o);
        if (halo > 0) {
          if (px >= 1 && px <= halo) put(halo - px, o);
          if (px >= W - 1 - halo && px <= W - 2) put(halo + 2 * (W - 1) - px, o);
        }
      }
    } else {
      bf16* dyb = reinterpret_cast<bf16*>(p.o1.ptr) + p.o1.off(n, row, x0) + c0;      // plain layouts: + pixel * C
      bf16* drb = p.has_o2 ? reinterpret_cast<bf16*>(p.o2.ptr) + p.o2.off(n, row, x0) + c0 : nullptr;
      const int pstep = 256 / cvs;
      int pl = tid / cvs;                               // pixel inside the segment
#pragma unroll 4
      for (int q = tid; q < nchunks; q += 256, pl += pstep) {
        float v[8], g[8], o[8];
        const uint4 gu = *reinterpret_cast<const uint4*>(sd + d_off(q));
        rp_unpack8(*reinterpret_cast<const uint4*>(sy + (size_t)q * 16), v);
        rp_unpack8(gu, g);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float dz = g[e] * rp_act_grad(sc[e] * v[e] + sh[e], act);
          o[e] = ba[e] * dz + bb[e] * v[e] + bc[e];
        }
        *reinterpret_cast<uint4*>(dyb + (long long)pl * C) = rp_pack8(o);
        if (p.has_o2) *reinterpret_cast<uint4*>(drb + (long long)pl * C) = gu;
      }
      if (x0 == 0) {
        rp_zero_halo(p.o1, n, row, cvs, tid);
        if (p.has_o2) rp_zero_halo(p.o2, n, row, cvs, tid);
      }
    }
  }

  if (MODE == RM_STATS || MODE == RM_BRED) {
    // block reduction over the pixel lanes (threads with the same channel group), through the drained ring
    __syncthreads();
    float2* red = reinterpret_cast<float2*>(rsm);      // [256 / cvs][C]
    const int pl = tid / cvs, PL = 256 / cvs;
#pragma unroll
    for (int e = 0; e < 8; ++e) red[pl * C + c0 + e] = make_float2(a0[e], a1[e]);
    __syncthreads();
    for (int c = tid; c < C; c += 256) {
      float t0 = 0.f, t1 = 0.f;
      for (int j = 0; j < PL; ++j) {
        const float2 v = red[j * C + c];
        t0 += v.x; t1 += v.y;
      }
      p.part[((long long)n * gridDim.x + blockIdx.x) * C + c] = make_float2(t0, t1);
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
static inline bool rowpipe_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DWC_ROWPIPE");
    on = e ? atoi(e) : 1;
  }
  return on != 0;
}

// geometry check + segmentation of a row into <= 16 KB pieces
static inline bool rowpipe_geom(const dwc_hbuf_t* y, int* nseg, int* segw, int* segbytes) {
  if (!rowpipe_enabled() || y->dtype != DWC_BF16 || y->layout != 0 || y->c % 8 != 0) return false;
  const int cvs = y->c / 8;
  if (cvs > 256 || 256 % cvs != 0) return false;
  const long long rowbytes = (long long)y->w * y->c * 2;
  if (rowbytes < 2048) return false;
  int k = 1;
  while (k <= y->w && !(y->w % k == 0 && rowbytes / k <= 16384)) ++k;
  if (k > y->w) return false;
  *nseg = k;
  *segw = y->w / k;
  *segbytes = (int)(rowbytes / k);
  return (*segbytes % 16) == 0;
}

// row_splits <= 0: pick the split count so that the whole grid is ONE wave of resident CTAs (a CTA streams many rows, so
// a partial second wave would cost as much as a full one)
template <int MODE>
static int rowpipe_launch(RowP& p, int nb, int row_splits, int n, cudaStream_t st) {
  p.d_planes = 0;
  p.dbytes = nb == 2 ? p.segbytes : 0;
  if (nb == 2 && p.d.layout == 1) {
    p.d_planes = 1;
    p.dbytes = p.d.wp * p.d.c * 2;
  }
  if (MODE == RM_BRED && p.d_fold) p.dbytes = p.d.wp * p.d.c * 2;      // whole padded rows of d
  const int stage_bytes = p.segbytes + p.dbytes;
  int stages = 103000 / stage_bytes;                   // 2 CTAs per SM next to ~11 KB of static shared memory
  if (stages > 4) stages = 4;
  if (stages < 2) stages = 2;
  p.stages = stages;
  size_t smem = (size_t)stages * stage_bytes;
  if (smem < 16384) smem = 16384;                      // the reduction scratch needs 256 * 8 float2
  static size_t attr = 0;
  if (smem > attr) {
    DWC_CUDA(cudaFuncSetAttribute(row_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  if (row_splits <= 0) {
    int per_sm = 1;                                    // resident CTAs per SM (registers and shared memory)
    DWC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, row_kernel<MODE>, 256, smem));
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    row_splits = (per_sm * dwc_num_sms()) / n;
    if (row_splits > p.y.h) row_splits = p.y.h;
    if (row_splits < 1) row_splits = 1;
  }
  DWC_CUDA(dwc_launch_pdl(row_kernel<MODE>, dim3(row_splits, n), dim3(256), smem, st, 1, p));
  DWC_LAUNCH_CHECK();
  return 0;
}
