// PARKED (csrc/experimental, built only with DWC_EXPERIMENTAL=1): measured in round 2 and slower than the three-kernel
// row-streaming backward it was meant to replace - 74 us against 49 us per 256x32x32 site at 48 images, 37 us against
// 24 us at 16 (profiles/r02d_norm_bwd_cluster.md).  186 KB of shared memory per CTA leaves one 8-warp CTA per SM and
// 128 co-resident CTAs (16 clusters), so 48 samples take three waves of a latency-bound CTA; the row-streaming kernels
// keep two CTAs per SM busy and find their second read of dout / y in L2.  Kept as a tested reference implementation
// of the cluster / distributed-shared-memory mechanics (DWC_NORM_CLUSTER=1 selects it in an experimental build).
#include "../common.cuh"

namespace {
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
}  // namespace

// ---------------------------------------------------------------------------------------------------
// One-pass backward of an InstanceNorm / AdaIN site with a whole sample held by a CLUSTER (round 2).
//
// The three-kernel backward (fold_halo, row_kernel<BRED>, row_kernel<BAPPLY>) reads dout and y twice and runs three
// launches on the critical path of backward.  Here a cluster of CL CTAs owns one sample: every CTA bulk-copies its
// H / CL image rows of y and the padded rows of dout (plus the reflected halo row that folds into them) into shared
// memory ONCE, folds the reflect-pad gradient while it forms dz, reduces sum(dz), sum(dz*y) per channel, exchanges
// the per-CTA partial sums through distributed shared memory (ld.shared::cluster), computes the coefficients of
// dy = a*dz + b*y + c, and writes dy (zero halo) and the residual gradient from shared memory: 3*E*2 bytes instead
// of 5*E*2 + a fold.  Sized for the 256-channel 32x32 residual blocks (32 of the 42 backward norm sites of a step):
// 4 rows per CTA, 64 KB of y + 85 KB of dout per CTA.
// ---------------------------------------------------------------------------------------------------
namespace {
constexpr int NB1_CL = 8;            // CTAs per cluster = per sample

struct Nb1P {
  HB dout, y, dy, dres;
  const float4* coef;
  const float* nweight;
  float* dweight;
  float* dbias;
  int kind, act, has_dres, rows;      // rows per CTA
};

__device__ __forceinline__ uint32_t nb1_mapa(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ float2 nb1_ld_cluster(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void nb1_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(NB1_CL, 1, 1) __launch_bounds__(256)
    norm_bwd_cluster_kernel(const __grid_constant__ Nb1P p) {
  extern __shared__ __align__(128) uint8_t nsm[];
  __shared__ uint64_t bar;
  __shared__ float2 part[512];                 // this CTA's per-channel {sum dz, sum dz*y}
  __shared__ float4 sbco[512];                 // {a, b, c, -} of dy = a*dz + b*y + c
  const int tid = threadIdx.x;
  const int rank = blockIdx.x;                 // == %cluster_ctarank (cluster dims (CL,1,1), grid (CL, N))
  const int n = blockIdx.y;
  const int H = p.y.h, W = p.y.w, C = p.y.c, R = p.rows;
  const int hd = p.dout.halo, Wp = W + 2 * hd;
  const int cvs = C >> 3, cv = tid % cvs, c0 = cv * 8, PL = 256 / cvs, pl = tid / cvs;
  const int r0 = rank * R;
  const uint32_t ybytes = (uint32_t)W * C * 2, dbytes = (uint32_t)Wp * C * 2;
  bf16* ybuf = reinterpret_cast<bf16*>(nsm);                                   // [R][W][C]
  bf16* dbuf = reinterpret_cast<bf16*>(nsm + (size_t)R * ybytes);              // [R + 2][Wp][C]: own rows, then reflections
  // reflections of halo rows that fold into this CTA's rows: padded row (hd - r) for interior row r in [1, hd],
  // padded row hd + 2(H-1) - r for r in [H-1-hd, H-2]  (hd <= 1: at most one each, rows 1 and H-2)
  const int top_row = (hd > 0 && 1 >= r0 && 1 < r0 + R) ? 1 : -1;              // interior row receiving padded row 0
  const int bot_row = (hd > 0 && H - 2 >= r0 && H - 2 < r0 + R) ? H - 2 : -1;  // interior row receiving padded row Hp-1
  const bf16* yb = reinterpret_cast<const bf16*>(p.y.ptr);
  const bf16* db = reinterpret_cast<const bf16*>(p.dout.ptr);

  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    const int nref = (top_row >= 0) + (bot_row >= 0);
    mbar_expect_tx(&bar, (uint32_t)R * (ybytes + dbytes) + (uint32_t)nref * dbytes);
    for (int i = 0; i < R; ++i) {
      bulk_load_1d(ybuf + (size_t)i * W * C, yb + p.y.off(n, r0 + i, 0), ybytes, &bar);
      bulk_load_1d(dbuf + (size_t)i * Wp * C, db + p.dout.off_padded(n, r0 + i + hd, 0), dbytes, &bar);
    }
    if (top_row >= 0) bulk_load_1d(dbuf + (size_t)R * Wp * C, db + p.dout.off_padded(n, 0, 0), dbytes, &bar);
    if (bot_row >= 0)
      bulk_load_1d(dbuf + (size_t)(R + 1) * Wp * C, db + p.dout.off_padded(n, H + 2 * hd - 1, 0), dbytes, &bar);
  }
  // forward coefficients of this thread's 8 channels (for the activation mask)
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float4 q = __ldg(p.coef + (long long)n * C + c0 + e);
    sc[e] = q.x; sh[e] = q.y;
  }
  __syncthreads();
  mbar_wait(&bar, 0);

  // ---- phase 1: fold the reflect-pad gradient (in shared memory, rounded to bf16 like the stored dres), reduce
  float a0[8], a1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a0[e] = a1[e] = 0.f;
  const int act = p.act;
  for (int i = 0; i < R; ++i) {
    const int row = r0 + i;
    const bf16* drow = dbuf + (size_t)i * Wp * C;
    const bf16* xrow = row == top_row ? dbuf + (size_t)R * Wp * C : (row == bot_row ? dbuf + (size_t)(R + 1) * Wp * C : nullptr);
    for (int px = pl; px < W; px += PL) {
      float g[8], t[8], v[8];
      rp_unpack8(*reinterpret_cast<const uint4*>(drow + (size_t)(px + hd) * C + c0), g);
      if (hd > 0) {
        const int xm = (px == 1) ? 0 : ((px == W - 2) ? Wp - 1 : -1);       // halo column that reflects onto px
        if (xm >= 0) {
          rp_unpack8(*reinterpret_cast<const uint4*>(drow + (size_t)xm * C + c0), t);
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] += t[e];
        }
        if (xrow) {
          rp_unpack8(*reinterpret_cast<const uint4*>(xrow + (size_t)(px + hd) * C + c0), t);
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] += t[e];
          if (xm >= 0) {
            rp_unpack8(*reinterpret_cast<const uint4*>(xrow + (size_t)xm * C + c0), t);
#pragma unroll
            for (int e = 0; e < 8; ++e) g[e] += t[e];
          }
        }
      }
      const uint4 gq = rp_pack8(g);                                          // the folded gradient as it is stored
      *reinterpret_cast<uint4*>(const_cast<bf16*>(drow) + (size_t)(px + hd) * C + c0) = gq;   // only its owner touches it
      rp_unpack8(gq, g);
      rp_unpack8(*reinterpret_cast<const uint4*>(ybuf + ((size_t)i * W + px) * C + c0), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float dz = g[e] * rp_act_grad(sc[e] * v[e] + sh[e], act);
        a0[e] += dz; a1[e] += dz * v[e];
      }
    }
  }
  // block reduction over the PL pixel lanes through the (now idle) tail of dbuf's reflection rows is not possible
  // (they may be in use): use a dedicated scratch behind the buffers
  float2* red = reinterpret_cast<float2*>(nsm + (size_t)R * ybytes + (size_t)(R + 2) * dbytes);   // [PL][C]
#pragma unroll
  for (int e = 0; e < 8; ++e) red[pl * C + c0 + e] = make_float2(a0[e], a1[e]);
  __syncthreads();
  for (int c = tid; c < C; c += 256) {
    float s0 = 0.f, s1 = 0.f;
    for (int j = 0; j < PL; ++j) {
      const float2 u = red[j * C + c];
      s0 += u.x; s1 += u.y;
    }
    part[c] = make_float2(s0, s1);
  }
  // ---- exchange the partial sums across the cluster, coefficients
  nb1_cluster_sync();
  const int hw = H * W;
  for (int c = tid; c < C; c += 256) {
    double S1 = 0, S2 = 0;
    for (int k = 0; k < NB1_CL; ++k) {                       // fixed order: deterministic
      const float2 u = nb1_ld_cluster(nb1_mapa(&part[c], (uint32_t)k));
      S1 += u.x; S2 += u.y;
    }
    const float4 q = p.coef[(long long)n * C + c];
    const double mean = q.z, rstd = q.w;
    const double w = p.kind == 2 ? (double)p.nweight[(long long)n * C + c] : 1.0;
    const double m1 = S1 / hw;
    const double m2 = rstd * (S2 / hw - mean * m1);
    if (p.kind == 2 && rank == 0) {
      p.dbias[(long long)n * C + c] = (float)S1;
      p.dweight[(long long)n * C + c] = (float)(rstd * (S2 - mean * S1));
    }
    const double a = w * rstd;
    const double b = -rstd * rstd * w * m2;
    const double cc = -rstd * w * m1 - b * mean;
    sbco[c] = make_float4((float)a, (float)b, (float)cc, 0.f);
  }
  __syncthreads();
  float ba[8], bb[8], bc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float4 q = sbco[c0 + e];
    ba[e] = q.x; bb[e] = q.y; bc[e] = q.z;
  }
  // ---- phase 2: apply from shared memory
  bf16* dyb = reinterpret_cast<bf16*>(p.dy.ptr);
  bf16* drb = reinterpret_cast<bf16*>(p.dres.ptr);
  for (int i = 0; i < R; ++i) {
    const int row = r0 + i;
    const bf16* drow = dbuf + (size_t)i * Wp * C;
    bf16* dyrow = dyb + p.dy.off(n, row, 0) + c0;
    bf16* drrow = p.has_dres ? drb + p.dres.off(n, row, 0) + c0 : nullptr;
    for (int px = pl; px < W; px += PL) {
      float g[8], v[8], o[8];
      const uint4 gq = *reinterpret_cast<const uint4*>(drow + (size_t)(px + hd) * C + c0);
      rp_unpack8(gq, g);
      rp_unpack8(*reinterpret_cast<const uint4*>(ybuf + ((size_t)i * W + px) * C + c0), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float dz = g[e] * rp_act_grad(sc[e] * v[e] + sh[e], act);
        o[e] = ba[e] * dz + bb[e] * v[e] + bc[e];
      }
      *reinterpret_cast<uint4*>(dyrow + (long long)px * C) = rp_pack8(o);
      if (p.has_dres) *reinterpret_cast<uint4*>(drrow + (long long)px * C) = gq;
    }
    rp_zero_halo(p.dy, n, row, cvs, tid);
    if (p.has_dres) rp_zero_halo(p.dres, n, row, cvs, tid);
  }
  nb1_cluster_sync();                          // nobody leaves while a peer may still read its partial sums
}
}  // namespace

// 1 if the site can run on the cluster kernel
extern "C" int dwc_post_bwd_cluster_ok(const dwc_hbuf_t* dout, const dwc_hbuf_t* y, int kind, const dwc_hbuf_t* dy,
                                       const dwc_hbuf_t* dres) {
  if (kind != 1 && kind != 2) return 0;
  if (y->dtype != DWC_BF16 || dout->dtype != DWC_BF16 || dy->dtype != DWC_BF16) return 0;
  if (y->layout != 0 || dout->layout != 0 || dy->layout != 0 || dout->halo > 1) return 0;
  if (dres && (dres->layout != 0 || dres->dtype != DWC_BF16 || dres->h != y->h || dres->w != y->w || dres->c != y->c)) return 0;
  if (y->c % 8 != 0 || y->c > 512 || 256 % (y->c / 8) != 0) return 0;
  if (y->h % NB1_CL != 0 || y->h < 2 * NB1_CL || y->w < 4) return 0;
  const long long rows = y->h / NB1_CL;
  const long long smem = rows * y->w * y->c * 2 + (rows + 2) * (long long)(y->w + 2 * dout->halo) * y->c * 2 +
                         (long long)(256 / (y->c / 8)) * y->c * 8;
  if (smem > 200 * 1024) return 0;
  if (((long long)y->w * y->c * 2) % 16 != 0) return 0;
  return 1;
}

extern "C" int dwc_post_bwd_cluster(const dwc_hbuf_t* dout, const dwc_hbuf_t* y, const float* coef, int kind, int act,
                                    const float* weight, float* dweight, float* dbias, const dwc_hbuf_t* dy,
                                    const dwc_hbuf_t* dres, dwc_stream_t stream) {
  DWC_CHECK(dwc_post_bwd_cluster_ok(dout, y, kind, dy, dres), "dwc_post_bwd_cluster: unsupported site geometry");
  DWC_CHECK(dout->n == y->n && dout->h == y->h && dout->w == y->w && dout->c == y->c && dy->n == y->n &&
                dy->h == y->h && dy->w == y->w && dy->c == y->c, "dwc_post_bwd_cluster: geometry mismatch");
  DWC_CHECK(kind != 2 || (weight && dweight && dbias), "dwc_post_bwd_cluster: AdaIN needs weight / dweight / dbias");
  Nb1P p;
  p.dout = HB(*dout); p.y = HB(*y); p.dy = HB(*dy); p.dres = dres ? HB(*dres) : HB(*dy);
  p.coef = reinterpret_cast<const float4*>(coef);
  p.nweight = weight; p.dweight = dweight; p.dbias = dbias;
  p.kind = kind; p.act = act; p.has_dres = dres != nullptr;
  p.rows = y->h / NB1_CL;
  const size_t smem = (size_t)p.rows * y->w * y->c * 2 + (size_t)(p.rows + 2) * (y->w + 2 * dout->halo) * y->c * 2 +
                      (size_t)(256 / (y->c / 8)) * y->c * 8;
  static size_t attr = 0;
  if (smem > attr) {
    DWC_CUDA(cudaFuncSetAttribute(norm_bwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  norm_bwd_cluster_kernel<<<dim3(NB1_CL, y->n), 256, smem, as_stream(stream)>>>(p);
  DWC_LAUNCH_CHECK();
  return 0;
}
