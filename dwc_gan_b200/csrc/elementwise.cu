// Vectorised NHWC element-wise / reduction kernels around the convolutions: per-(n,c) statistics,
// normalisation (+activation, +residual, +reflect padding) forward and backward, bilinear 2x upsampling,
// image <-> haloed-buffer conversion, decoder heads, attention blend, global average pooling.
// All are HBM-bound: 16-byte (bf16) / 32-byte (fp32) accesses along the channel axis, fp32 math.
#include "common.cuh"
#include <stdlib.h>

#define DISPATCH_T(dtype, ...)                     \
  do {                                             \
    if ((dtype) == DWC_F32) {                      \
      typedef float T;                             \
      __VA_ARGS__;                                 \
    } else {                                       \
      typedef bf16 T;                              \
      __VA_ARGS__;                                 \
    }                                              \
  } while (0)

__device__ __forceinline__ float act_fwd(float z, int act) {
  if (act == 1) return z > 0.f ? z : 0.f;
  if (act == 2) return z > 0.f ? z : 0.1f * z;
  return z;
}
__device__ __forceinline__ float act_grad(float z, int act) {
  if (act == 1) return z > 0.f ? 1.f : 0.f;
  if (act == 2) return z > 0.f ? 1.f : 0.1f;
  return 1.f;
}

// sum of dout over all padded positions that reflect onto interior (y,x): 8 channels starting at c0
template <typename T>
__device__ __forceinline__ void fold_read8(const HB& d, int n, int y, int x, int c0, float* f) {
  const T* base = reinterpret_cast<const T*>(d.ptr);
  int ys[3], xs[3], ny = 0, nx = 0;
  ys[ny++] = y;
  xs[nx++] = x;
  if (d.refl > 0) {
    if (y >= 1 && y <= d.refl) ys[ny++] = -y;
    if (y >= d.h - 1 - d.refl && y <= d.h - 2) ys[ny++] = 2 * (d.h - 1) - y;
    if (x >= 1 && x <= d.refl) xs[nx++] = -x;
    if (x >= d.w - 1 - d.refl && x <= d.w - 2) xs[nx++] = 2 * (d.w - 1) - x;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = 0.f;
  for (int i = 0; i < ny; ++i)
    for (int j = 0; j < nx; ++j) {
      float t[8];
      Vec8<T>::load(base + d.off(n, ys[i], xs[j]) + c0, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += t[e];
    }
}

// scalar version (any channel count)
template <typename T>
__device__ __forceinline__ float fold_read1(const HB& d, int n, int y, int x, int c) {
  const T* base = reinterpret_cast<const T*>(d.ptr);
  int ys[3], xs[3], ny = 0, nx = 0;
  ys[ny++] = y;
  xs[nx++] = x;
  if (d.refl > 0) {
    if (y >= 1 && y <= d.refl) ys[ny++] = -y;
    if (y >= d.h - 1 - d.refl && y <= d.h - 2) ys[ny++] = 2 * (d.h - 1) - y;
    if (x >= 1 && x <= d.refl) xs[nx++] = -x;
    if (x >= d.w - 1 - d.refl && x <= d.w - 2) xs[nx++] = 2 * (d.w - 1) - x;
  }
  float s = 0.f;
  for (int i = 0; i < ny; ++i)
    for (int j = 0; j < nx; ++j) s += to_f<T>(base[d.off(n, ys[i], xs[j]) + c]);
  return s;
}

// ---------------------------------------------------------------------------------------------------
// per-(n,c) statistics.  grid (C/32, splits, N); block 256 = 4 channel-vectors x 64 pixel lanes.
// MODE 0: {sum y, sum y^2}.  MODE 1: {sum dz, sum dz*y} (backward).
// ---------------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
    nc_reduce_kernel(HB y, HB dout, const float4* __restrict__ coef, int act, int splits, float2* __restrict__ out) {
  __shared__ float s0[64][33], s1[64][33];
  const int cv = threadIdx.x & 3, pl = threadIdx.x >> 2;
  const int c0 = blockIdx.x * 32 + cv * 8;
  const int split = blockIdx.y, n = blockIdx.z;
  const int hw = y.h * y.w;
  const int per = (hw + splits - 1) / splits;
  const int p_begin = split * per, p_end = min(hw, p_begin + per);
  const T* yb = reinterpret_cast<const T*>(y.ptr);
  float a0[8], a1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a0[e] = a1[e] = 0.f;
  float sc[8], sh[8];
  if (MODE == 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (coef) {
        float4 q = coef[(long long)n * y.c + c0 + e];
        sc[e] = q.x; sh[e] = q.y;
      } else { sc[e] = 1.f; sh[e] = 0.f; }
    }
  }
  if (c0 < y.c) {
    for (int p = p_begin + pl; p < p_end; p += 64) {
      int py = p / y.w, px = p - py * y.w;
      float v[8];
      Vec8<T>::load(yb + y.off(n, py, px) + c0, v);
      if (MODE == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { a0[e] += v[e]; a1[e] += v[e] * v[e]; }
      } else {
        float g[8];
        fold_read8<T>(dout, n, py, px, c0, g);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float dz = g[e] * act_grad(sc[e] * v[e] + sh[e], act);
          a0[e] += dz; a1[e] += dz * v[e];
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) { s0[pl][cv * 8 + e] = a0[e]; s1[pl][cv * 8 + e] = a1[e]; }
  __syncthreads();
  if (threadIdx.x < 32) {
    float t0 = 0.f, t1 = 0.f;
    for (int i = 0; i < 64; ++i) { t0 += s0[i][threadIdx.x]; t1 += s1[i][threadIdx.x]; }
    int c = blockIdx.x * 32 + threadIdx.x;
    if (c < y.c) out[((long long)n * splits + split) * y.c + c] = make_float2(t0, t1);
  }
}

// ---------------------------------------------------------------------------------------------------
// statistics -> scale/shift   (one block per sample)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_d(double v, double* sm) {
  const int tid = threadIdx.x;
  sm[tid] = v;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (tid < o) sm[tid] += sm[tid + o];
    __syncthreads();
  }
  double r = sm[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(256)
    norm_finalize_kernel(int kind, const float2* __restrict__ stats, int splits, int C, int hw, float eps,
                         const float* __restrict__ weight, const float* __restrict__ bias, float4* __restrict__ coef) {
  __shared__ double sm[256];
  const int n = blockIdx.x;
  if (kind == 1 || kind == 2) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double s = 0, q = 0;
      for (int k = 0; k < splits; ++k) {
        float2 v = stats[((long long)n * splits + k) * C + c];
        s += v.x; q += v.y;
      }
      double mean = s / hw;
      double var = q / hw - mean * mean;
      if (var < 0) var = 0;
      float rstd = (float)(1.0 / sqrt(var + (double)eps));
      float w = kind == 2 ? weight[(long long)n * C + c] : 1.f;
      float b = kind == 2 ? bias[(long long)n * C + c] : 0.f;
      coef[(long long)n * C + c] = make_float4(w * rstd, b - (float)mean * rstd * w, (float)mean, rstd);
    }
  } else if (kind == 3) {
    double s = 0, q = 0;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      for (int k = 0; k < splits; ++k) {
        float2 v = stats[((long long)n * splits + k) * C + c];
        s += v.x; q += v.y;
      }
    s = block_sum_d(s, sm);
    q = block_sum_d(q, sm);
    const double M = (double)C * hw;
    double mean = s / M;
    double var = (q - M * mean * mean) / (M - 1.0);
    if (var < 0) var = 0;
    double sd = sqrt(var);
    float inv = (float)(1.0 / (sd + (double)eps));
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float g = weight[c], b = bias[c];
      coef[(long long)n * C + c] = make_float4(g * inv, b - (float)mean * g * inv, (float)mean, inv);
    }
  } else {
    for (int c = threadIdx.x; c < C; c += blockDim.x) coef[(long long)n * C + c] = make_float4(1.f, 0.f, 0.f, 1.f);
  }
}

extern "C" int dwc_norm_finalize(int kind, const float* stats, int splits, int n, int c, int hw, float eps,
                                 const float* weight, const float* bias, float* coef, dwc_stream_t stream) {
  DWC_CHECK(kind >= 0 && kind <= 3, "dwc_norm_finalize: bad kind");
  norm_finalize_kernel<<<n, 256, 0, as_stream(stream)>>>(kind, reinterpret_cast<const float2*>(stats), splits, c, hw,
                                                          eps, weight, bias, reinterpret_cast<float4*>(coef));
  DWC_LAUNCH_CHECK();
  return 0;
}

// backward coefficients: dy = a*dz + b*y + c.   IN/AdaIN: one block per sample.  LN: single block.
__global__ void __launch_bounds__(256)
    norm_bwd_finalize_kernel(int kind, const float2* __restrict__ red, int splits, const float4* __restrict__ coef,
                             int N, int C, int hw, float eps, const float* __restrict__ weight,
                             float* __restrict__ dweight, float* __restrict__ dbias, float4* __restrict__ bco) {
  __shared__ double sm[256];
  if (kind == 1 || kind == 2) {
    const int n = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double S1 = 0, S2 = 0;
      for (int k = 0; k < splits; ++k) {
        float2 v = red[((long long)n * splits + k) * C + c];
        S1 += v.x; S2 += v.y;
      }
      float4 q = coef[(long long)n * C + c];
      double mean = q.z, rstd = q.w;
      double w = kind == 2 ? (double)weight[(long long)n * C + c] : 1.0;
      double m1 = S1 / hw;
      double m2 = rstd * (S2 / hw - mean * m1);
      if (kind == 2) {
        dbias[(long long)n * C + c] = (float)S1;
        dweight[(long long)n * C + c] = (float)(rstd * (S2 - mean * S1));
      }
      double a = w * rstd;
      double b = -rstd * rstd * w * m2;
      double cc = -rstd * w * m1 - b * mean;
      bco[(long long)n * C + c] = make_float4((float)a, (float)b, (float)cc, 0.f);
    }
  } else if (kind == 3) {
    // per-sample coefficients: one block per sample
    const double M = (double)C * hw;
    const int n = blockIdx.x;
    double g1 = 0, g2 = 0;
    const double mean = coef[(long long)n * C].z;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double S1 = 0, S2 = 0;
      for (int k = 0; k < splits; ++k) {
        float2 v = red[((long long)n * splits + k) * C + c];
        S1 += v.x; S2 += v.y;
      }
      double g = weight[c];
      g1 += g * S1;
      g2 += g * (S2 - mean * S1);
    }
    g1 = block_sum_d(g1, sm);
    g2 = block_sum_d(g2, sm);
    const double inv = coef[(long long)n * C].w;
    const double sd = 1.0 / inv - (double)eps;
    const double K = sd > 0 ? g2 * inv * inv / ((M - 1.0) * sd) : 0.0;
    const double b = -K;
    const double cc = -g1 * inv / M + K * mean;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      bco[(long long)n * C + c] = make_float4((float)(weight[c] * inv), (float)b, (float)cc, 0.f);
  }
}

// LayerNorm parameter gradients (accumulated): one warp per channel, lanes over (sample, split) pairs, fixed
// shuffle-tree summation order (deterministic)
__global__ void ln_param_grad_kernel(const float2* __restrict__ red, int splits, const float4* __restrict__ coef, int N,
                                     int C, float* __restrict__ dweight, float* __restrict__ dbias) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  float dg = 0.f, db = 0.f;
  for (int i = lane; i < N * splits; i += 32) {
    const int n = i / splits;
    const float2 v = red[(long long)i * C + c];
    const float4 q = coef[(long long)n * C + c];
    dg += q.w * (v.y - q.z * v.x);
    db += v.x;
  }
  dg = warp_sum(dg);
  db = warp_sum(db);
  if (lane == 0) {
    dweight[c] += dg;
    dbias[c] += db;
  }
}

extern "C" int dwc_norm_bwd_finalize(int kind, const float* red, int splits, const float* coef, int n, int c, int hw,
                                     float eps, const float* weight, float* dweight, float* dbias, float* bco,
                                     dwc_stream_t stream) {
  DWC_CHECK(kind >= 1 && kind <= 3, "dwc_norm_bwd_finalize: bad kind");
  norm_bwd_finalize_kernel<<<n, 256, 0, as_stream(stream)>>>(
      kind, reinterpret_cast<const float2*>(red), splits, reinterpret_cast<const float4*>(coef), n, c, hw, eps, weight,
      dweight, dbias, reinterpret_cast<float4*>(bco));
  DWC_LAUNCH_CHECK();
  if (kind == 3) {
    ln_param_grad_kernel<<<cdiv((long long)c * 32, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float2*>(red), splits,
                                                                   reinterpret_cast<const float4*>(coef), n, c, dweight,
                                                                   dbias);
    DWC_LAUNCH_CHECK();
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// forward post pass: out(padded) = reflect_pad(act(scale*y+shift) + res)
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
    post_fwd_kernel(HB y, const float4* __restrict__ coef, int act, HB res, int has_res, HB out) {
  const int cvs = out.c >> 3;
  const long long total = out.padded_pixels() * cvs;
  const T* yb = reinterpret_cast<const T*>(y.ptr);
  const T* rb = reinterpret_cast<const T*>(res.ptr);
  T* ob = reinterpret_cast<T*>(out.ptr);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % cvs);
    long long pix = i / cvs;
    int X = (int)(pix % out.wp);
    long long r = pix / out.wp;
    int Y = (int)(r % out.hp);
    int n = (int)(r / out.hp);
    int iy = reflect_idx(Y - out.halo, out.h), ix = reflect_idx(X - out.halo, out.w);
    const int c0 = cv * 8;
    float v[8];
    Vec8<T>::load(yb + y.off(n, iy, ix) + c0, v);
    if (coef) {
      const float4* q = coef + (long long)n * out.c + c0;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = act_fwd(q[e].x * v[e] + q[e].y, act);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = act_fwd(v[e], act);
    }
    if (has_res) {
      float rr[8];
      Vec8<T>::load(rb + res.off(n, iy, ix) + c0, rr);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += rr[e];
    }
    Vec8<T>::store(ob + out.off_padded(n, Y, X) + c0, v);
  }
}

// Per-sample variant (grid.y = sample, 256 %% (C/8) == 0): every thread keeps ONE 8-channel group for the whole launch,
// so its scale / shift coefficients are loaded once instead of once per pixel.
template <typename T>
__global__ void __launch_bounds__(256)
    post_fwd_ps_kernel(HB y, const float4* __restrict__ coef, int act, HB res, int has_res, HB out) {
  const int cvs = out.c >> 3;
  const int cv = threadIdx.x % cvs, pl = threadIdx.x / cvs, PL = 256 / cvs;
  const int n = blockIdx.y, c0 = cv * 8;
  const T* yb = reinterpret_cast<const T*>(y.ptr);
  const T* rb = reinterpret_cast<const T*>(res.ptr);
  T* ob = reinterpret_cast<T*>(out.ptr);
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (coef) {
      const float4 q = coef[(long long)n * out.c + c0 + e];
      sc[e] = q.x; sh[e] = q.y;
    } else { sc[e] = 1.f; sh[e] = 0.f; }
  }
  const int npix = out.hp * out.wp;
  for (int p = blockIdx.x * PL + pl; p < npix; p += gridDim.x * PL) {
    const int Y = p / out.wp, X = p - Y * out.wp;
    const int iy = reflect_idx(Y - out.halo, out.h), ix = reflect_idx(X - out.halo, out.w);
    float v[8];
    Vec8<T>::load(yb + y.off(n, iy, ix) + c0, v);
    if (coef) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = act_fwd(sc[e] * v[e] + sh[e], act);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = act_fwd(v[e], act);
    }
    if (has_res) {
      float rr[8];
      Vec8<T>::load(rb + res.off(n, iy, ix) + c0, rr);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += rr[e];
    }
    Vec8<T>::store(ob + out.off_padded(n, Y, X) + c0, v);
  }
}

// ---------------------------------------------------------------------------------------------------
// bf16 fast paths of the HBM-bound passes.  One thread owns one 8-channel group of a sample for the whole launch
// (coefficients in registers) and works on PF pixels at a time: all 16-byte loads of the batch are issued before the
// first use, which is what keeps enough bytes in flight to approach the HBM roofline (a single load per loop
// iteration leaves the memory system latency-bound at ~1.5 TB/s).
// ---------------------------------------------------------------------------------------------------
constexpr int PF = 4;

__device__ __forceinline__ uint4 ld16(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bool has_reflection(int i, int size, int halo) {
  return halo > 0 && ((i >= 1 && i <= halo) || (i >= size - 1 - halo && i <= size - 2));
}

#include "rowpipe.cuh"

__global__ void __launch_bounds__(256)
    post_fwd_fast_kernel(HB y, const float4* __restrict__ coef, int act, HB res, int has_res, HB out) {
  const int cvs = out.c >> 3;
  const int cv = threadIdx.x % cvs, pl = threadIdx.x / cvs, PL = 256 / cvs;
  const int n = blockIdx.y, c0 = cv * 8;
  const bf16* yb = reinterpret_cast<const bf16*>(y.ptr);
  const bf16* rb = reinterpret_cast<const bf16*>(res.ptr);
  bf16* ob = reinterpret_cast<bf16*>(out.ptr);
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (coef) {
      const float4 q = coef[(long long)n * out.c + c0 + e];
      sc[e] = q.x; sh[e] = q.y;
    } else { sc[e] = 1.f; sh[e] = 0.f; }
  }
  const int npix = out.hp * out.wp;
  const int stride = gridDim.x * PL;
  for (int p0 = blockIdx.x * PL + pl; p0 < npix; p0 += PF * stride) {
    uint4 vy[PF], vr[PF];
    long long oo[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int p = p0 + u * stride;
      oo[u] = -1;
      if (p < npix) {
        const int Y = p / out.wp, X = p - Y * out.wp;
        const int iy = reflect_idx(Y - out.halo, out.h), ix = reflect_idx(X - out.halo, out.w);
        vy[u] = ld16(yb + y.off(n, iy, ix) + c0);
        if (has_res) vr[u] = ld16(rb + res.off(n, iy, ix) + c0);
        oo[u] = out.off_padded(n, Y, X) + c0;
      }
    }
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      if (oo[u] < 0) continue;
      float v[8];
      unpack8(vy[u], v);
      if (coef) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = act_fwd(sc[e] * v[e] + sh[e], act);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = act_fwd(v[e], act);
      }
      if (has_res) {
        float rr[8];
        unpack8(vr[u], rr);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += rr[e];
      }
      Vec8<bf16>::store(ob + oo[u], v);
    }
  }
}

__global__ void __launch_bounds__(256, 2)
    post_bwd_apply_fast_kernel(HB dout, HB y, const float4* __restrict__ coef, const float4* __restrict__ bco, int act,
                               HB dy, HB dres, int has_dres) {
  const int hmax = max(dy.halo, has_dres ? dres.halo : 0);
  const int HP = y.h + 2 * hmax, WP = y.w + 2 * hmax;
  const int cvs = y.c >> 3;
  const int cv = threadIdx.x % cvs, pl = threadIdx.x / cvs, PL = 256 / cvs;
  const int n = blockIdx.y, c0 = cv * 8;
  const bf16* yb = reinterpret_cast<const bf16*>(y.ptr);
  const bf16* db = reinterpret_cast<const bf16*>(dout.ptr);
  bf16* dyb = reinterpret_cast<bf16*>(dy.ptr);
  bf16* drb = reinterpret_cast<bf16*>(dres.ptr);
  float sc[8], sh[8], ba[8], bb[8], bc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (coef) {
      const float4 q = coef[(long long)n * y.c + c0 + e];
      sc[e] = q.x; sh[e] = q.y;
    } else { sc[e] = 1.f; sh[e] = 0.f; }
    if (bco) {
      const float4 q = bco[(long long)n * y.c + c0 + e];
      ba[e] = q.x; bb[e] = q.y; bc[e] = q.z;
    } else { ba[e] = 1.f; bb[e] = 0.f; bc[e] = 0.f; }
  }
  const int npix = HP * WP;
  const int stride = gridDim.x * PL;
  for (int p0 = blockIdx.x * PL + pl; p0 < npix; p0 += PF * stride) {
    uint4 vy[PF], vd[PF];
    int iys[PF], ixs[PF], kind[PF];          // kind: 0 skip, 1 outside the interior, 2 interior, 3 interior + reflections
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int p = p0 + u * stride;
      kind[u] = 0;
      if (p < npix) {
        const int Y = p / WP, X = p - Y * WP;
        const int iy = Y - hmax, ix = X - hmax;
        iys[u] = iy; ixs[u] = ix;
        kind[u] = 1;
        if (iy >= 0 && iy < y.h && ix >= 0 && ix < y.w) {
          vy[u] = ld16(yb + y.off(n, iy, ix) + c0);
          if (has_reflection(iy, dout.h, dout.refl) || has_reflection(ix, dout.w, dout.refl)) {
            kind[u] = 3;
          } else {
            kind[u] = 2;
            vd[u] = ld16(db + dout.off(n, iy, ix) + c0);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      if (kind[u] == 0) continue;
      const int iy = iys[u], ix = ixs[u];
      float g[8], o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = o[e] = 0.f;
      if (kind[u] >= 2) {
        if (kind[u] == 3) fold_read8<bf16>(dout, n, iy, ix, c0, g);
        else unpack8(vd[u], g);
        float v[8];
        unpack8(vy[u], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float z = coef ? sc[e] * v[e] + sh[e] : v[e];
          const float dz = g[e] * act_grad(z, act);
          o[e] = bco ? ba[e] * dz + bb[e] * v[e] + bc[e] : dz;
        }
      }
      {
        const int py = iy + dy.halo, px = ix + dy.halo;
        if (py >= 0 && py < dy.hp && px >= 0 && px < dy.wp) Vec8<bf16>::store(dyb + dy.off_padded(n, py, px) + c0, o);
      }
      if (has_dres) {
        const int py = iy + dres.halo, px = ix + dres.halo;
        if (py >= 0 && py < dres.hp && px >= 0 && px < dres.wp)
          Vec8<bf16>::store(drb + dres.off_padded(n, py, px) + c0, g);
      }
    }
  }
}

// per-(n,c) sums, bf16.  grid (splits, N); a block covers ALL channels of its pixel range: thread = (8-channel group,
// pixel lane), so a pixel is one contiguous C*2-byte read; PF pixels per thread are in flight at a time.
template <int MODE>
__global__ void __launch_bounds__(256, 2)
    nc_reduce_fast_kernel(HB y, HB dout, const float4* __restrict__ coef, int act, int splits, float2* __restrict__ out) {
  __shared__ float2 red[2048];                       // [pixel lane][channel], 256/cvs * C == 2048 entries
  const int cvs = y.c >> 3;
  const int cv = threadIdx.x % cvs, pl = threadIdx.x / cvs, PL = 256 / cvs;
  const int c0 = cv * 8;
  const int split = blockIdx.x, n = blockIdx.y;
  const int hw = y.h * y.w;
  const int per = (hw + splits - 1) / splits;
  const int p_begin = split * per, p_end = min(hw, p_begin + per);
  const bf16* yb = reinterpret_cast<const bf16*>(y.ptr);
  const bf16* db = reinterpret_cast<const bf16*>(dout.ptr);
  float a0[8], a1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a0[e] = a1[e] = 0.f;
  float sc[8], sh[8];
  if (MODE == 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (coef) {
        float4 q = coef[(long long)n * y.c + c0 + e];
        sc[e] = q.x; sh[e] = q.y;
      } else { sc[e] = 1.f; sh[e] = 0.f; }
    }
  }
  for (int pb = p_begin + pl; pb < p_end; pb += PL * PF) {
    uint4 vy[PF], vd[PF];
    int kind[PF], pys[PF], pxs[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int p = pb + u * PL;
      kind[u] = 0;
      if (p < p_end) {
        const int py = p / y.w, px = p - py * y.w;
        pys[u] = py; pxs[u] = px;
        vy[u] = ld16(yb + y.off(n, py, px) + c0);
        kind[u] = 2;
        if (MODE == 1) {
          if (has_reflection(py, dout.h, dout.refl) || has_reflection(px, dout.w, dout.refl)) kind[u] = 3;
          else vd[u] = ld16(db + dout.off(n, py, px) + c0);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      if (kind[u] == 0) continue;
      float v[8];
      unpack8(vy[u], v);
      if (MODE == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { a0[e] += v[e]; a1[e] += v[e] * v[e]; }
      } else {
        float g[8];
        if (kind[u] == 3) fold_read8<bf16>(dout, n, pys[u], pxs[u], c0, g);
        else unpack8(vd[u], g);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float dz = g[e] * act_grad(sc[e] * v[e] + sh[e], act);
          a0[e] += dz; a1[e] += dz * v[e];
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[pl * y.c + c0 + e] = make_float2(a0[e], a1[e]);
  __syncthreads();
  for (int c = threadIdx.x; c < y.c; c += 256) {
    float t0 = 0.f, t1 = 0.f;
    for (int i = 0; i < PL; ++i) {
      const float2 v = red[i * y.c + c];
      t0 += v.x; t1 += v.y;
    }
    out[((long long)n * splits + split) * y.c + c] = make_float2(t0, t1);
  }
}

// grid for the per-sample kernels: enough blocks per sample to fill the GPU ~8 CTAs deep
static inline dim3 ps_grid(int npix, int cvs, int n) {
  const int PL = 256 / cvs;
  int bx = (npix + PL - 1) / PL;
  int cap = (dwc_num_sms() * 8 + n - 1) / n;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  return dim3(bx < 1 ? 1 : bx, n);
}
static inline bool ps_ok(int c) { return c % 8 == 0 && (c / 8) <= 256 && 256 % (c / 8) == 0; }

static inline int ew_grid(long long total) {
  long long b = (total + 255) / 256;
  long long cap = (long long)dwc_num_sms() * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

extern "C" int dwc_post_fwd(const dwc_hbuf_t* y, const float* coef, int act, const dwc_hbuf_t* res,
                            const dwc_hbuf_t* out, dwc_stream_t stream) {
  DWC_CHECK(y->c % 8 == 0, "dwc_post_fwd: needs C %% 8 == 0");
  DWC_CHECK(y->n == out->n && y->h == out->h && y->w == out->w && y->c == out->c && y->dtype == out->dtype,
            "dwc_post_fwd: geometry mismatch");
  DWC_CHECK(out->layout == 0 || ((out->h + 2 * out->halo) % 2 == 0 && (out->w + 2 * out->halo) % 2 == 0),
            "dwc_post_fwd: plane layout needs even padded extent");
  HB hy(*y), ho(*out), hr = res ? HB(*res) : HB(*y);
  long long total = ho.padded_pixels() * (out->c / 8);
  {
    RowP rp{};
    if (rowpipe_geom(y, &rp.nseg, &rp.segw, &rp.segbytes) && out->dtype == DWC_BF16 && (!res || res->layout == 0) &&
        out->halo <= y->h - 1 && out->halo <= y->w - 1) {
      rp.y = hy; rp.d = hr; rp.o1 = ho; rp.o2 = ho;
      rp.coef = reinterpret_cast<const float4*>(coef); rp.bco = nullptr; rp.part = nullptr;
      rp.act = act; rp.has_d = res != nullptr; rp.has_o2 = 0;
      return rowpipe_launch<RM_FWD>(rp, res ? 2 : 1, 0, y->n, as_stream(stream));
    }
  }
  if (ps_ok(out->c)) {
    if (y->dtype == DWC_BF16)
      post_fwd_fast_kernel<<<ps_grid(ho.hp * ho.wp, out->c / 8, out->n), 256, 0, as_stream(stream)>>>(
          hy, reinterpret_cast<const float4*>(coef), act, hr, res != nullptr, ho);
    else
      post_fwd_ps_kernel<float><<<ps_grid(ho.hp * ho.wp, out->c / 8, out->n), 256, 0, as_stream(stream)>>>(
          hy, reinterpret_cast<const float4*>(coef), act, hr, res != nullptr, ho);
  } else {
    DISPATCH_T(y->dtype, (post_fwd_kernel<T><<<ew_grid(total), 256, 0, as_stream(stream)>>>(
                             hy, reinterpret_cast<const float4*>(coef), act, hr, res != nullptr, ho)));
  }
  DWC_LAUNCH_CHECK();
  return 0;
}

// statistics -> coefficients -> normalise / activation / residual / pad: the coefficient step runs inside the
// row-streaming kernel when the site is eligible (no dwc_norm_finalize launch), else as the separate passes
extern "C" int dwc_post_fwd_norm(const dwc_hbuf_t* y, int kind, const float* stats, int splits, float eps,
                                 const float* weight, const float* bias, int act, const dwc_hbuf_t* res,
                                 const dwc_hbuf_t* out, float* coef, dwc_stream_t stream) {
  DWC_CHECK(kind >= 1 && kind <= 3, "dwc_post_fwd_norm: bad kind");
  DWC_CHECK(y->n == out->n && y->h == out->h && y->w == out->w && y->c == out->c && y->dtype == out->dtype,
            "dwc_post_fwd_norm: geometry mismatch");
  RowP rp{};
  if (rowpipe_geom(y, &rp.nseg, &rp.segw, &rp.segbytes) && y->c <= 512 && (!res || res->layout == 0) &&
      out->halo <= y->h - 1 && out->halo <= y->w - 1 &&
      (out->layout == 0 || ((out->h + 2 * out->halo) % 2 == 0 && (out->w + 2 * out->halo) % 2 == 0))) {
    HB hy(*y), ho(*out), hr = res ? HB(*res) : HB(*y);
    rp.y = hy; rp.d = hr; rp.o1 = ho; rp.o2 = ho;
    rp.act = act; rp.has_d = res != nullptr; rp.has_o2 = 0;
    rp.nstats = reinterpret_cast<const float2*>(stats); rp.nweight = weight; rp.nbias = bias;
    rp.coef_out = reinterpret_cast<float4*>(coef);
    rp.nkind = kind; rp.nsplits = splits; rp.nhw = y->h * y->w; rp.neps = eps;
    return rowpipe_launch<RM_FWD>(rp, res ? 2 : 1, 0, y->n, as_stream(stream));
  }
  if (dwc_norm_finalize(kind, stats, splits, y->n, y->c, y->h * y->w, eps, weight, bias, coef, stream)) return 1;
  return dwc_post_fwd(y, coef, act, res, out, stream);
}

// ---------------------------------------------------------------------------------------------------
// backward post pass: dy = a*dz + b*y + c (zero halo), dres = fold(dout) (zero halo)
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
    post_bwd_apply_kernel(HB dout, HB y, const float4* __restrict__ coef, const float4* __restrict__ bco, int act, HB dy,
                          HB dres, int has_dres) {
  const int hmax = max(dy.halo, has_dres ? dres.halo : 0);
  const int HP = y.h + 2 * hmax, WP = y.w + 2 * hmax;
  const int cvs = y.c >> 3;
  const long long total = (long long)y.n * HP * WP * cvs;
  const T* yb = reinterpret_cast<const T*>(y.ptr);
  T* dyb = reinterpret_cast<T*>(dy.ptr);
  T* drb = reinterpret_cast<T*>(dres.ptr);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % cvs);
    long long pix = i / cvs;
    int X = (int)(pix % WP);
    long long r = pix / WP;
    int Y = (int)(r % HP);
    int n = (int)(r / HP);
    const int iy = Y - hmax, ix = X - hmax;     // interior coordinates (may be outside)
    const int c0 = cv * 8;
    const bool interior = iy >= 0 && iy < y.h && ix >= 0 && ix < y.w;
    float g[8], o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = o[e] = 0.f;
    if (interior) {
      fold_read8<T>(dout, n, iy, ix, c0, g);
      float v[8];
      Vec8<T>::load(yb + y.off(n, iy, ix) + c0, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float z = v[e], dz = g[e];
        if (coef) {
          float4 q = coef[(long long)n * y.c + c0 + e];
          z = q.x * v[e] + q.y;
        }
        dz *= act_grad(z, act);
        if (bco) {
          float4 q = bco[(long long)n * y.c + c0 + e];
          o[e] = q.x * dz + q.y * v[e] + q.z;
        } else {
          o[e] = dz;
        }
      }
    }
    // dy extent
    {
      int py = iy + dy.halo, px = ix + dy.halo;
      if (py >= 0 && py < dy.hp && px >= 0 && px < dy.wp) Vec8<T>::store(dyb + dy.off_padded(n, py, px) + c0, o);
    }
    if (has_dres) {
      int py = iy + dres.halo, px = ix + dres.halo;
      if (py >= 0 && py < dres.hp && px >= 0 && px < dres.wp) Vec8<T>::store(drb + dres.off_padded(n, py, px) + c0, g);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    post_bwd_apply_ps_kernel(HB dout, HB y, const float4* __restrict__ coef, const float4* __restrict__ bco, int act,
                             HB dy, HB dres, int has_dres) {
  const int hmax = max(dy.halo, has_dres ? dres.halo : 0);
  const int HP = y.h + 2 * hmax, WP = y.w + 2 * hmax;
  const int cvs = y.c >> 3;
  const int cv = threadIdx.x % cvs, pl = threadIdx.x / cvs, PL = 256 / cvs;
  const int n = blockIdx.y, c0 = cv * 8;
  const T* yb = reinterpret_cast<const T*>(y.ptr);
  T* dyb = reinterpret_cast<T*>(dy.ptr);
  T* drb = reinterpret_cast<T*>(dres.ptr);
  float sc[8], sh[8], ba[8], bb[8], bc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (coef) {
      const float4 q = coef[(long long)n * y.c + c0 + e];
      sc[e] = q.x; sh[e] = q.y;
    } else { sc[e] = 1.f; sh[e] = 0.f; }
    if (bco) {
      const float4 q = bco[(long long)n * y.c + c0 + e];
      ba[e] = q.x; bb[e] = q.y; bc[e] = q.z;
    } else { ba[e] = 1.f; bb[e] = 0.f; bc[e] = 0.f; }
  }
  const int npix = HP * WP;
  for (int p = blockIdx.x * PL + pl; p < npix; p += gridDim.x * PL) {
    const int Y = p / WP, X = p - Y * WP;
    const int iy = Y - hmax, ix = X - hmax;     // interior coordinates (may be outside)
    const bool interior = iy >= 0 && iy < y.h && ix >= 0 && ix < y.w;
    float g[8], o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = o[e] = 0.f;
    if (interior) {
      fold_read8<T>(dout, n, iy, ix, c0, g);
      float v[8];
      Vec8<T>::load(yb + y.off(n, iy, ix) + c0, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float z = coef ? sc[e] * v[e] + sh[e] : v[e];
        const float dz = g[e] * act_grad(z, act);
        o[e] = bco ? ba[e] * dz + bb[e] * v[e] + bc[e] : dz;
      }
    }
    {
      int py = iy + dy.halo, px = ix + dy.halo;
      if (py >= 0 && py < dy.hp && px >= 0 && px < dy.wp) Vec8<T>::store(dyb + dy.off_padded(n, py, px) + c0, o);
    }
    if (has_dres) {
      int py = iy + dres.halo, px = ix + dres.halo;
      if (py >= 0 && py < dres.hp && px >= 0 && px < dres.wp) Vec8<T>::store(drb + dres.off_padded(n, py, px) + c0, g);
    }
  }
}

extern "C" int dwc_post_bwd_apply(const dwc_hbuf_t* dout, const dwc_hbuf_t* y, const float* coef, const float* bco,
                                  int act, const dwc_hbuf_t* dy, const dwc_hbuf_t* dres, int prefolded,
                                  dwc_stream_t stream) {
  DWC_CHECK(y->c % 8 == 0 && y->layout == 0 && dy->layout == 0, "dwc_post_bwd_apply: needs C %% 8 == 0, plain y/dy");
  DWC_CHECK(dout->n == y->n && dout->h == y->h && dout->w == y->w && dout->c == y->c, "dwc_post_bwd_apply: geometry mismatch");
  HB hd(*dout), hy(*y), hdy(*dy), hr = dres ? HB(*dres) : HB(*dy);
  if (prefolded) hd.refl = 0;
  {
    RowP rp{};
    if (rowpipe_geom(y, &rp.nseg, &rp.segw, &rp.segbytes) && (prefolded || dout->halo == 0) && dy->dtype == DWC_BF16 &&
        dout->dtype == DWC_BF16 && (!dres || (dres->layout == 0 && dres->dtype == DWC_BF16)) &&
        (dout->layout == 0 || rp.nseg == 1)) {
      rp.y = hy; rp.d = hd; rp.o1 = hdy; rp.o2 = hr;
      rp.coef = reinterpret_cast<const float4*>(coef); rp.bco = reinterpret_cast<const float4*>(bco); rp.part = nullptr;
      rp.act = act; rp.has_d = 1; rp.has_o2 = dres != nullptr;
      return rowpipe_launch<RM_BAPPLY>(rp, 2, 0, y->n, as_stream(stream));
    }
  }
  int hmax = dy->halo;
  if (dres && dres->halo > hmax) hmax = dres->halo;
  long long total = (long long)y->n * (y->h + 2 * hmax) * (y->w + 2 * hmax) * (y->c / 8);
  if (ps_ok(y->c)) {
    const dim3 g = ps_grid((y->h + 2 * hmax) * (y->w + 2 * hmax), y->c / 8, y->n);
    if (y->dtype == DWC_BF16)
      post_bwd_apply_fast_kernel<<<g, 256, 0, as_stream(stream)>>>(hd, hy, reinterpret_cast<const float4*>(coef),
                                                                   reinterpret_cast<const float4*>(bco), act, hdy, hr,
                                                                   dres != nullptr);
    else
      post_bwd_apply_ps_kernel<float><<<g, 256, 0, as_stream(stream)>>>(hd, hy, reinterpret_cast<const float4*>(coef),
                                                                        reinterpret_cast<const float4*>(bco), act, hdy,
                                                                        hr, dres != nullptr);
  } else {
    DISPATCH_T(y->dtype, (post_bwd_apply_kernel<T><<<ew_grid(total), 256, 0, as_stream(stream)>>>(
                             hd, hy, reinterpret_cast<const float4*>(coef), reinterpret_cast<const float4*>(bco), act,
                             hdy, hr, dres != nullptr)));
  }
  DWC_LAUNCH_CHECK();
  return 0;
}

// reductions -> coefficients (+ parameter gradients) -> apply: the coefficient step runs inside the row-streaming
// kernel when the site is eligible (no dwc_norm_bwd_finalize launch), else as the separate passes (bco = scratch [N,C,4])
extern "C" int dwc_post_bwd_apply_norm(const dwc_hbuf_t* dout, const dwc_hbuf_t* y, const float* coef, int kind,
                                       const float* red, int splits, float eps, const float* weight, float* dweight,
                                       float* dbias, float* bco, int act, const dwc_hbuf_t* dy, const dwc_hbuf_t* dres,
                                       int prefolded, dwc_stream_t stream) {
  DWC_CHECK(kind >= 1 && kind <= 3, "dwc_post_bwd_apply_norm: bad kind");
  DWC_CHECK(y->c % 8 == 0 && y->layout == 0 && dy->layout == 0, "dwc_post_bwd_apply_norm: needs C %% 8 == 0, plain y/dy");
  DWC_CHECK(dout->n == y->n && dout->h == y->h && dout->w == y->w && dout->c == y->c,
            "dwc_post_bwd_apply_norm: geometry mismatch");
  RowP rp{};
  if (rowpipe_geom(y, &rp.nseg, &rp.segw, &rp.segbytes) && y->c <= 512 && (prefolded || dout->halo == 0) &&
      dy->dtype == DWC_BF16 && dout->dtype == DWC_BF16 && (!dres || (dres->layout == 0 && dres->dtype == DWC_BF16)) &&
      (dout->layout == 0 || rp.nseg == 1)) {
    HB hd(*dout), hy(*y), hdy(*dy), hr = dres ? HB(*dres) : HB(*dy);
    hd.refl = 0;
    rp.y = hy; rp.d = hd; rp.o1 = hdy; rp.o2 = hr;
    rp.coef = reinterpret_cast<const float4*>(coef);
    rp.act = act; rp.has_d = 1; rp.has_o2 = dres != nullptr;
    rp.nstats = reinterpret_cast<const float2*>(red); rp.nweight = weight;
    rp.dweight = dweight; rp.dbias = dbias;
    rp.nkind = kind; rp.nsplits = splits; rp.nhw = y->h * y->w; rp.neps = eps;
    if (rowpipe_launch<RM_BAPPLY>(rp, 2, 0, y->n, as_stream(stream))) return 1;
    if (kind == 3) {
      ln_param_grad_kernel<<<cdiv((long long)y->c * 32, 256), 256, 0, as_stream(stream)>>>(
          reinterpret_cast<const float2*>(red), splits, reinterpret_cast<const float4*>(coef), y->n, y->c, dweight, dbias);
      DWC_LAUNCH_CHECK();
    }
    return 0;
  }
  if (dwc_norm_bwd_finalize(kind, red, splits, coef, y->n, y->c, y->h * y->w, eps, weight, dweight, dbias, bco, stream))
    return 1;
  return dwc_post_bwd_apply(dout, y, coef, bco, act, dy, dres, prefolded, stream);
}

// ---------------------------------------------------------------------------------------------------
// Fused InstanceNorm / AdaIN sites for small feature maps (H*W <= 1536: the 32x32 residual blocks, i.e. most norm
// sites of the network).  One CTA owns a (sample, channel group) slab that fits in shared memory, so the site is ONE
// kernel and ONE read of its inputs: forward = statistics + coefficients + normalise/activation/residual/reflect pad;
// backward = reflect fold + per-channel reductions + coefficients + apply.  (The three-pass kernels above read the
// tensors twice and need two more launches for the coefficients.)
// ---------------------------------------------------------------------------------------------------
template <int CG>
__global__ void __launch_bounds__(256)
    post_fused_fwd_kernel(HB y, int kind, const float* __restrict__ weight, const float* __restrict__ bias, float eps,
                          int act, HB res, int has_res, HB out, float4* __restrict__ coef_out) {
  extern __shared__ __align__(16) uint8_t fsm[];
  constexpr int CVS = CG / 8, PL = 256 / CVS;
  const int hw = y.h * y.w;
  bf16* slab = reinterpret_cast<bf16*>(fsm);                         // [hw][CG]
  float2* red = reinterpret_cast<float2*>(fsm + (size_t)hw * CG * 2);  // [PL][CG]
  float2* scoef = red + PL * CG;                                      // [CG] scale, shift
  const int cv = threadIdx.x % CVS, pl = threadIdx.x / CVS;
  const int n = blockIdx.y, cbase = blockIdx.x * CG, c0 = cbase + cv * 8;
  const bf16* yb = reinterpret_cast<const bf16*>(y.ptr);
  float a0[8], a1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a0[e] = a1[e] = 0.f;
  for (int pb = pl; pb < hw; pb += PL * PF) {
    uint4 v[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int p = pb + u * PL;
      if (p < hw) {
        const int py = p / y.w, px = p - py * y.w;
        v[u] = ld16(yb + y.off(n, py, px) + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int p = pb + u * PL;
      if (p < hw) {
        *reinterpret_cast<uint4*>(slab + (size_t)p * CG + cv * 8) = v[u];
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) { a0[e] += f[e]; a1[e] += f[e] * f[e]; }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[pl * CG + cv * 8 + e] = make_float2(a0[e], a1[e]);
  __syncthreads();
  if (threadIdx.x < CG) {
    double s = 0, q = 0;
    for (int i = 0; i < PL; ++i) {
      const float2 t = red[i * CG + threadIdx.x];
      s += t.x; q += t.y;
    }
    const int c = cbase + threadIdx.x;
    const double mean = s / hw;
    double var = q / hw - mean * mean;
    if (var < 0) var = 0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float w = kind == 2 ? weight[(long long)n * y.c + c] : 1.f;
    const float b = kind == 2 ? bias[(long long)n * y.c + c] : 0.f;
    const float4 cf = make_float4(w * rstd, b - (float)mean * rstd * w, (float)mean, rstd);
    coef_out[(long long)n * y.c + c] = cf;
    scoef[threadIdx.x] = make_float2(cf.x, cf.y);
  }
  __syncthreads();
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float2 t = scoef[cv * 8 + e];
    sc[e] = t.x; sh[e] = t.y;
  }
  const bf16* rb = reinterpret_cast<const bf16*>(res.ptr);
  bf16* ob = reinterpret_cast<bf16*>(out.ptr);
  const int npix = out.hp * out.wp;
  for (int p = pl; p < npix; p += PL) {
    const int Y = p / out.wp, X = p - Y * out.wp;
    const int iy = reflect_idx(Y - out.halo, out.h), ix = reflect_idx(X - out.halo, out.w);
    float v[8];
    unpack8(*reinterpret_cast<const uint4*>(slab + (size_t)(iy * y.w + ix) * CG + cv * 8), v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = act_fwd(sc[e] * v[e] + sh[e], act);
    if (has_res) {
      float rr[8];
      unpack8(ld16(rb + res.off(n, iy, ix) + c0), rr);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += rr[e];
    }
    Vec8<bf16>::store(ob + out.off_padded(n, Y, X) + c0, v);
  }
}

template <int CG>
__global__ void __launch_bounds__(256)
    post_fused_bwd_kernel(HB dout, HB y, const float4* __restrict__ coef, int kind, int act,
                          const float* __restrict__ weight, float* __restrict__ dweight, float* __restrict__ dbias,
                          HB dy, HB dres, int has_dres) {
  extern __shared__ __align__(16) uint8_t fsm[];
  constexpr int CVS = CG / 8, PL = 256 / CVS;
  const int hw = y.h * y.w;
  bf16* ys = reinterpret_cast<bf16*>(fsm);                            // [hw][CG]
  bf16* gs = ys + (size_t)hw * CG;                                    // [hw][CG] folded gradient
  float2* red = reinterpret_cast<float2*>(fsm + (size_t)hw * CG * 4);  // [PL][CG]
  float4* sbco = reinterpret_cast<float4*>(red + PL * CG);            // [CG] a, b, c
  const int cv = threadIdx.x % CVS, pl = threadIdx.x / CVS;
  const int n = blockIdx.y, cbase = blockIdx.x * CG, c0 = cbase + cv * 8;
  const bf16* yb = reinterpret_cast<const bf16*>(y.ptr);
  const bf16* db = reinterpret_cast<const bf16*>(dout.ptr);
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float4 q = coef[(long long)n * y.c + c0 + e];
    sc[e] = q.x; sh[e] = q.y;
  }
  float a0[8], a1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a0[e] = a1[e] = 0.f;
  for (int pb = pl; pb < hw; pb += PL * PF) {
    uint4 vy[PF], vd[PF];
    int kind_u[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int p = pb + u * PL;
      kind_u[u] = 0;
      if (p < hw) {
        const int py = p / y.w, px = p - py * y.w;
        vy[u] = ld16(yb + y.off(n, py, px) + c0);
        if (has_reflection(py, dout.h, dout.refl) || has_reflection(px, dout.w, dout.refl)) kind_u[u] = 3;
        else {
          kind_u[u] = 2;
          vd[u] = ld16(db + dout.off(n, py, px) + c0);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      if (kind_u[u] == 0) continue;
      const int p = pb + u * PL;
      float v[8], g[8];
      unpack8(vy[u], v);
      if (kind_u[u] == 3) {
        const int py = p / y.w, px = p - py * y.w;
        fold_read8<bf16>(dout, n, py, px, c0, g);
        Vec8<bf16>::store(gs + (size_t)p * CG + cv * 8, g);
        Vec8<bf16>::load(gs + (size_t)p * CG + cv * 8, g);          // the folded gradient is a bf16 tensor (dres)
      } else {
        unpack8(vd[u], g);
        *reinterpret_cast<uint4*>(gs + (size_t)p * CG + cv * 8) = vd[u];
      }
      *reinterpret_cast<uint4*>(ys + (size_t)p * CG + cv * 8) = vy[u];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float dz = g[e] * act_grad(sc[e] * v[e] + sh[e], act);
        a0[e] += dz; a1[e] += dz * v[e];
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[pl * CG + cv * 8 + e] = make_float2(a0[e], a1[e]);
  __syncthreads();
  if (threadIdx.x < CG) {
    double S1 = 0, S2 = 0;
    for (int i = 0; i < PL; ++i) {
      const float2 t = red[i * CG + threadIdx.x];
      S1 += t.x; S2 += t.y;
    }
    const int c = cbase + threadIdx.x;
    const float4 q = coef[(long long)n * y.c + c];
    const double mean = q.z, rstd = q.w;
    const double w = kind == 2 ? (double)weight[(long long)n * y.c + c] : 1.0;
    const double m1 = S1 / hw;
    const double m2 = rstd * (S2 / hw - mean * m1);
    if (kind == 2) {
      dbias[(long long)n * y.c + c] = (float)S1;
      dweight[(long long)n * y.c + c] = (float)(rstd * (S2 - mean * S1));
    }
    const double a = w * rstd;
    const double b = -rstd * rstd * w * m2;
    const double cc = -rstd * w * m1 - b * mean;
    sbco[threadIdx.x] = make_float4((float)a, (float)b, (float)cc, 0.f);
  }
  __syncthreads();
  float ba[8], bb[8], bc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float4 t = sbco[cv * 8 + e];
    ba[e] = t.x; bb[e] = t.y; bc[e] = t.z;
  }
  const int hmax = max(dy.halo, has_dres ? dres.halo : 0);
  const int HP = y.h + 2 * hmax, WP = y.w + 2 * hmax;
  bf16* dyb = reinterpret_cast<bf16*>(dy.ptr);
  bf16* drb = reinterpret_cast<bf16*>(dres.ptr);
  for (int p = pl; p < HP * WP; p += PL) {
    const int Y = p / WP, X = p - Y * WP;
    const int iy = Y - hmax, ix = X - hmax;
    float g[8], o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = o[e] = 0.f;
    if (iy >= 0 && iy < y.h && ix >= 0 && ix < y.w) {
      const size_t so = (size_t)(iy * y.w + ix) * CG + cv * 8;
      float v[8];
      unpack8(*reinterpret_cast<const uint4*>(ys + so), v);
      unpack8(*reinterpret_cast<const uint4*>(gs + so), g);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float dz = g[e] * act_grad(sc[e] * v[e] + sh[e], act);
        o[e] = ba[e] * dz + bb[e] * v[e] + bc[e];
      }
    }
    {
      const int py = iy + dy.halo, px = ix + dy.halo;
      if (py >= 0 && py < dy.hp && px >= 0 && px < dy.wp) Vec8<bf16>::store(dyb + dy.off_padded(n, py, px) + c0, o);
    }
    if (has_dres) {
      const int py = iy + dres.halo, px = ix + dres.halo;
      if (py >= 0 && py < dres.hp && px >= 0 && px < dres.wp)
        Vec8<bf16>::store(drb + dres.off_padded(n, py, px) + c0, g);
    }
  }
}

constexpr int FUSED_FWD_CG = 32, FUSED_BWD_CG = 16;
constexpr int FUSED_MAX_HW = 1536;

extern "C" int dwc_post_fused_ok(int c, int hw, int dtype) {
  return dtype == DWC_BF16 && hw <= FUSED_MAX_HW && c % FUSED_FWD_CG == 0;
}

extern "C" int dwc_post_fused_fwd(const dwc_hbuf_t* y, int kind, const float* weight, const float* bias, float eps,
                                  int act, const dwc_hbuf_t* res, const dwc_hbuf_t* out, float* coef,
                                  dwc_stream_t stream) {
  DWC_CHECK((kind == 1 || kind == 2) && dwc_post_fused_ok(y->c, y->h * y->w, y->dtype) && y->layout == 0,
            "dwc_post_fused_fwd: unsupported site (kind %d, C %d, %dx%d)", kind, y->c, y->h, y->w);
  DWC_CHECK(y->n == out->n && y->h == out->h && y->w == out->w && y->c == out->c && out->dtype == DWC_BF16,
            "dwc_post_fused_fwd: geometry mismatch");
  HB hy(*y), ho(*out), hr = res ? HB(*res) : HB(*y);
  constexpr int CG = FUSED_FWD_CG;
  const size_t smem = (size_t)y->h * y->w * CG * 2 + (256 / (CG / 8)) * CG * sizeof(float2) + CG * sizeof(float2);
  static size_t attr = 0;
  if (smem > attr) {
    DWC_CUDA(cudaFuncSetAttribute(post_fused_fwd_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  post_fused_fwd_kernel<CG><<<dim3(y->c / CG, y->n), 256, smem, as_stream(stream)>>>(
      hy, kind, weight, bias, eps, act, hr, res != nullptr, ho, reinterpret_cast<float4*>(coef));
  DWC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dwc_post_fused_bwd(const dwc_hbuf_t* dout, const dwc_hbuf_t* y, const float* coef, int kind, int act,
                                  const float* weight, float* dweight, float* dbias, const dwc_hbuf_t* dy,
                                  const dwc_hbuf_t* dres, dwc_stream_t stream) {
  DWC_CHECK((kind == 1 || kind == 2) && dwc_post_fused_ok(y->c, y->h * y->w, y->dtype) && y->layout == 0 &&
                dy->layout == 0,
            "dwc_post_fused_bwd: unsupported site (kind %d, C %d, %dx%d)", kind, y->c, y->h, y->w);
  DWC_CHECK(dout->n == y->n && dout->h == y->h && dout->w == y->w && dout->c == y->c, "dwc_post_fused_bwd: geometry mismatch");
  HB hd(*dout), hy(*y), hdy(*dy), hr = dres ? HB(*dres) : HB(*dy);
  constexpr int CG = FUSED_BWD_CG;
  const size_t smem = (size_t)y->h * y->w * CG * 4 + (256 / (CG / 8)) * CG * sizeof(float2) + CG * sizeof(float4);
  static size_t attr = 0;
  if (smem > attr) {
    DWC_CUDA(cudaFuncSetAttribute(post_fused_bwd_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  post_fused_bwd_kernel<CG><<<dim3(y->c / CG, y->n), 256, smem, as_stream(stream)>>>(
      hd, hy, reinterpret_cast<const float4*>(coef), kind, act, weight, dweight, dbias, hdy, hr, dres != nullptr);
  DWC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// In-place fold of a reflect halo's gradient into the interior: interior pixels whose mirror images lie in the halo
// (a band of `halo` rows / columns next to each edge) take the sum of their reflections.  Only those band pixels are
// visited; afterwards the backward passes stream the interior without any gather (refl = 0).
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fold_halo_kernel(HB d) {
  pdl_prologue();
  const int cvs = d.c >> 3;
  const int nb = 2 * d.halo;                       // band rows (and band columns)
  const int row_part = nb * d.w, col_part = (d.h - nb) * nb;
  const long long total = (long long)d.n * (row_part + col_part) * cvs;
  T* base = reinterpret_cast<T*>(d.ptr);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvs);
    long long r = i / cvs;
    const int j = (int)(r % (row_part + col_part));
    const int n = (int)(r / (row_part + col_part));
    int y, x;
    if (j < row_part) {
      const int rr = j / d.w;
      x = j - rr * d.w;
      y = rr < d.halo ? 1 + rr : (d.h - 1 - d.halo) + (rr - d.halo);
    } else {
      const int jj = j - row_part;
      const int k = jj / nb, cc = jj - k * nb;
      y = k == 0 ? 0 : (k == d.h - nb - 1 ? d.h - 1 : d.halo + k);
      x = cc < d.halo ? 1 + cc : (d.w - 1 - d.halo) + (cc - d.halo);
    }
    float g[8];
    fold_read8<T>(d, n, y, x, cv * 8, g);
    Vec8<T>::store(base + d.off(n, y, x) + cv * 8, g);
  }
}

extern "C" int dwc_fold_halo(const dwc_hbuf_t* d, dwc_stream_t stream) {
  DWC_CHECK(d->c % 8 == 0, "dwc_fold_halo: needs C %% 8 == 0");
  DWC_CHECK(d->halo > 0 && d->h >= 2 * d->halo + 2 && d->w >= 2 * d->halo + 2,
            "dwc_fold_halo: image %dx%d too small for halo %d", d->h, d->w, d->halo);
  HB hd(*d);
  const int nb = 2 * d->halo;
  long long total = (long long)d->n * (nb * d->w + (d->h - nb) * nb) * (d->c / 8);
  if (d->dtype == DWC_F32)
    DWC_CUDA(dwc_launch_pdl(fold_halo_kernel<float>, dim3(ew_grid(total)), dim3(256), 0, as_stream(stream), 1, hd));
  else
    DWC_CUDA(dwc_launch_pdl(fold_halo_kernel<bf16>, dim3(ew_grid(total)), dim3(256), 0, as_stream(stream), 1, hd));
  DWC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// bilinear 2x upsample (align_corners = False) + reflect pad
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void up_taps(int o, int size_in, int* i0, int* i1, float* w0, float* w1) {
  float src = (o + 0.5f) * 0.5f - 0.5f;
  if (src < 0.f) src = 0.f;
  int a = (int)src;
  int b = a + 1 < size_in ? a + 1 : size_in - 1;
  float l1 = src - a;
  *i0 = a; *i1 = b; *w0 = 1.f - l1; *w1 = l1;
}

template <typename T>
__global__ void __launch_bounds__(256) upsample_pad_fwd_kernel(HB x, HB out) {
  const int cvs = out.c >> 3;
  const long long total = out.padded_pixels() * cvs;
  const T* xb = reinterpret_cast<const T*>(x.ptr);
  T* ob = reinterpret_cast<T*>(out.ptr);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % cvs);
    long long pix = i / cvs;
    int X = (int)(pix % out.wp);
    long long r = pix / out.wp;
    int Y = (int)(r % out.hp);
    int n = (int)(r / out.hp);
    int oy = reflect_idx(Y - out.halo, out.h), ox = reflect_idx(X - out.halo, out.w);
    int y0, y1, x0, x1;
    float wy0, wy1, wx0, wx1;
    up_taps(oy, x.h, &y0, &y1, &wy0, &wy1);
    up_taps(ox, x.w, &x0, &x1, &wx0, &wx1);
    const int c0 = cv * 8;
    float a[8], b[8], c[8], d[8], o[8];
    Vec8<T>::load(xb + x.off(n, y0, x0) + c0, a);
    Vec8<T>::load(xb + x.off(n, y0, x1) + c0, b);
    Vec8<T>::load(xb + x.off(n, y1, x0) + c0, c);
    Vec8<T>::load(xb + x.off(n, y1, x1) + c0, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = wy0 * (wx0 * a[e] + wx1 * b[e]) + wy1 * (wx0 * c[e] + wx1 * d[e]);
    Vec8<T>::store(ob + out.off_padded(n, Y, X) + c0, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) upsample_pad_bwd_kernel(HB dout, HB dx) {
  // dx interior (h,w); dout interior (2h,2w) with reflect halo to fold
  const int cvs = dx.c >> 3;
  const long long total = dx.padded_pixels() * cvs;
  T* dxb = reinterpret_cast<T*>(dx.ptr);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % cvs);
    long long pix = i / cvs;
    int X = (int)(pix % dx.wp);
    long long r = pix / dx.wp;
    int Y = (int)(r % dx.hp);
    int n = (int)(r / dx.hp);
    const int iy = Y - dx.halo, ix = X - dx.halo;
    const int c0 = cv * 8;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    if (iy >= 0 && iy < dx.h && ix >= 0 && ix < dx.w) {
      for (int oy = max(0, 2 * iy - 1); oy <= min(dout.h - 1, 2 * iy + 2); ++oy) {
        int y0, y1;
        float wy0, wy1;
        up_taps(oy, dx.h, &y0, &y1, &wy0, &wy1);
        float wy = (y0 == iy ? wy0 : 0.f) + (y1 == iy ? wy1 : 0.f);
        if (wy == 0.f) continue;
        for (int ox = max(0, 2 * ix - 1); ox <= min(dout.w - 1, 2 * ix + 2); ++ox) {
          int x0, x1;
          float wx0, wx1;
          up_taps(ox, dx.w, &x0, &x1, &wx0, &wx1);
          float wx = (x0 == ix ? wx0 : 0.f) + (x1 == ix ? wx1 : 0.f);
          if (wx == 0.f) continue;
          float g[8];
          fold_read8<T>(dout, n, oy, ox, c0, g);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += wy * wx * g[e];
        }
      }
    }
    Vec8<T>::store(dxb + dx.off_padded(n, Y, X) + c0, acc);
  }
}

// bf16 fast path: grid.y = sample, PF output pixels per thread in flight (4 x PF 16-byte loads before the first use)
__global__ void __launch_bounds__(256) upsample_pad_fwd_fast_kernel(HB x, HB out) {
  const int cvs = out.c >> 3;
  const int cv = threadIdx.x % cvs, pl = threadIdx.x / cvs, PL = 256 / cvs;
  const int n = blockIdx.y, c0 = cv * 8;
  const bf16* xb = reinterpret_cast<const bf16*>(x.ptr);
  bf16* ob = reinterpret_cast<bf16*>(out.ptr);
  const int npix = out.hp * out.wp;
  const int stride = gridDim.x * PL;
  constexpr int UF = 2;
  for (int p0 = blockIdx.x * PL + pl; p0 < npix; p0 += UF * stride) {
    uint4 va[UF], vb[UF], vc[UF], vd[UF];
    float wy0[UF], wy1[UF], wx0[UF], wx1[UF];
    long long oo[UF];
#pragma unroll
    for (int u = 0; u < UF; ++u) {
      const int p = p0 + u * stride;
      oo[u] = -1;
      if (p < npix) {
        const int Y = p / out.wp, X = p - Y * out.wp;
        const int oy = reflect_idx(Y - out.halo, out.h), ox = reflect_idx(X - out.halo, out.w);
        int y0, y1, x0, x1;
        up_taps(oy, x.h, &y0, &y1, &wy0[u], &wy1[u]);
        up_taps(ox, x.w, &x0, &x1, &wx0[u], &wx1[u]);
        va[u] = ld16(xb + x.off(n, y0, x0) + c0);
        vb[u] = ld16(xb + x.off(n, y0, x1) + c0);
        vc[u] = ld16(xb + x.off(n, y1, x0) + c0);
        vd[u] = ld16(xb + x.off(n, y1, x1) + c0);
        oo[u] = out.off_padded(n, Y, X) + c0;
      }
    }
#pragma unroll
    for (int u = 0; u < UF; ++u) {
      if (oo[u] < 0) continue;
      float a[8], b[8], c[8], d[8], o[8];
      unpack8(va[u], a); unpack8(vb[u], b); unpack8(vc[u], c); unpack8(vd[u], d);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        o[e] = wy0[u] * (wx0[u] * a[e] + wx1[u] * b[e]) + wy1[u] * (wx0[u] * c[e] + wx1[u] * d[e]);
      Vec8<bf16>::store(ob + oo[u], o);
    }
  }
}

// bf16 fast path of the backward: gradient already folded into the interior (refl == 0); one input pixel per thread
// iteration, its 4 x 4 candidate output pixels loaded together (16 x 16-byte loads in flight)
__global__ void __launch_bounds__(256) upsample_pad_bwd_fast_kernel(HB dout, HB dx) {
  const int cvs = dx.c >> 3;
  const int cv = threadIdx.x % cvs, pl = threadIdx.x / cvs, PL = 256 / cvs;
  const int n = blockIdx.y, c0 = cv * 8;
  const bf16* db = reinterpret_cast<const bf16*>(dout.ptr);
  bf16* dxb = reinterpret_cast<bf16*>(dx.ptr);
  const int npix = dx.hp * dx.wp;
  for (int p = blockIdx.x * PL + pl; p < npix; p += gridDim.x * PL) {
    const int Y = p / dx.wp, X = p - Y * dx.wp;
    const int iy = Y - dx.halo, ix = X - dx.halo;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    if (iy >= 0 && iy < dx.h && ix >= 0 && ix < dx.w) {
      float wy[4], wx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int oy = 2 * iy - 1 + j, ox = 2 * ix - 1 + j;
        wy[j] = 0.f;
        wx[j] = 0.f;
        if (oy >= 0 && oy < dout.h) {
          int a, b;
          float w0, w1;
          up_taps(oy, dx.h, &a, &b, &w0, &w1);
          wy[j] = (a == iy ? w0 : 0.f) + (b == iy ? w1 : 0.f);
        }
        if (ox >= 0 && ox < dout.w) {
          int a, b;
          float w0, w1;
          up_taps(ox, dx.w, &a, &b, &w0, &w1);
          wx[j] = (a == ix ? w0 : 0.f) + (b == ix ? w1 : 0.f);
        }
      }
      uint4 v[16];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int oy = min(max(2 * iy - 1 + j, 0), dout.h - 1), ox = min(max(2 * ix - 1 + i, 0), dout.w - 1);
          v[j * 4 + i] = ld16(db + dout.off(n, oy, ox) + c0);      // clamped address; weight is 0 outside
        }
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float w = wy[j] * wx[i];
          float g[8];
          unpack8(v[j * 4 + i], g);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, g[e], acc[e]);
        }
    }
    Vec8<bf16>::store(dxb + dx.off_padded(n, Y, X) + c0, acc);
  }
}

// Row-streaming forward (bf16): input rows arrive by 1-D bulk copies in a shared-memory ring; a CTA owns a range of
// input rows and produces the two output rows between each consecutive pair (plus the image's first / last row), one
// thread per (pair of neighbouring input columns, 8 channels): 4 shared-memory reads give a 2 x 2 block of outputs.
struct UpP {
  HB x, out;
  int stages, rowbytes;
};

__global__ void __launch_bounds__(256) row_up_fwd_kernel(const __grid_constant__ UpP p) {
  extern __shared__ __align__(128) uint8_t usm[];
  __shared__ uint64_t full[8];
  const int tid = threadIdx.x, n = blockIdx.y;
  const int H = p.x.h, W = p.x.w, C = p.x.c, cvs = C >> 3, S = p.stages;
  const int r0 = (int)(((long long)blockIdx.x * H) / gridDim.x), r1 = (int)(((long long)(blockIdx.x + 1) * H) / gridDim.x);
  if (r1 <= r0) return;
  const int lo = max(r0 - 1, 0), hi = min(r1, H - 1), cnt = hi - lo + 1;
  const bf16* xb = reinterpret_cast<const bf16*>(p.x.ptr);
  bf16* ob = reinterpret_cast<bf16*>(p.out.ptr);
  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  int issued = 0;                                       // thread 0 only
  auto issue_upto = [&](int last) {
    while (issued < cnt && issued <= last) {
      const int s = issued % S;
      mbar_expect_tx(&full[s], (uint32_t)p.rowbytes);
      bulk_load_1d(usm + (size_t)s * p.rowbytes, xb + p.x.off(n, lo + issued, 0), (uint32_t)p.rowbytes, &full[s]);
      ++issued;
    }
  };
  if (tid == 0) issue_upto(S - 1);
  const int halo = p.out.halo, HO = 2 * H, WO = 2 * W;
  int waited = -1;                                      // highest unit whose barrier this thread has passed
  for (int a = r0 - 1; a <= r1 - 1; ++a) {
    const int ia = max(a, 0), ib = min(a + 1, H - 1);
    __syncthreads();                                    // the previous pair is done: units below ia - lo are dead
    if (tid == 0) issue_upto(ia - lo + S - 1);
    for (int u = waited + 1; u <= ib - lo; ++u) mbar_wait(&full[u % S], (uint32_t)((u / S) & 1));
    waited = max(waited, ib - lo);
    const uint8_t* rowA = usm + (size_t)((ia - lo) % S) * p.rowbytes;
    const uint8_t* rowB = usm + (size_t)((ib - lo) % S) * p.rowbytes;
    // output rows 2a+1 (weights .75 / .25 on rows a, a+1) and 2a+2 (.25 / .75; the image's first row takes row 0 alone)
    const int oyk[2] = {2 * a + 1, 2 * a + 2};
    bool vk[2];
    int Yk[2][3];                                       // primary padded row and up to two reflected copies (-1 = none)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int oy = oyk[k];
      vk[k] = oy >= 2 * r0 && oy < 2 * r1 && oy >= 0 && oy < HO;
      Yk[k][0] = oy + halo;
      Yk[k][1] = (halo > 0 && oy >= 1 && oy <= halo) ? halo - oy : -1;
      Yk[k][2] = (halo > 0 && oy >= HO - 1 - halo && oy <= HO - 2) ? halo + 2 * (HO - 1) - oy : -1;
    }
    const float wyk[2][2] = {{0.75f, 0.25f}, {a < 0 ? 1.f : 0.25f, a < 0 ? 0.f : 0.75f}};
    for (int t = tid; t < (W + 1) * cvs; t += 256) {
      const int jj = t / cvs - 1, cv = t - (jj + 1) * cvs;
      const int ja = max(jj, 0), jb = min(jj + 1, W - 1);
      float va[8], vb[8], vc[8], vd[8];
      rp_unpack8(*reinterpret_cast<const uint4*>(rowA + ((size_t)ja * cvs + cv) * 16), va);
      rp_unpack8(*reinterpret_cast<const uint4*>(rowA + ((size_t)jb * cvs + cv) * 16), vb);
      rp_unpack8(*reinterpret_cast<const uint4*>(rowB + ((size_t)ja * cvs + cv) * 16), vc);
      rp_unpack8(*reinterpret_cast<const uint4*>(rowB + ((size_t)jb * cvs + cv) * 16), vd);
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        // output columns 2jj+1 (.75 / .25 on columns jj, jj+1) and 2jj+2 (.25 / .75; column 0 takes column 0 alone)
        const int ox = 2 * jj + 1 + m;
        if (ox < 0 || ox >= WO) continue;
        const float wx0 = m == 0 ? 0.75f : (jj < 0 ? 1.f : 0.25f), wx1 = m == 0 ? 0.25f : (jj < 0 ? 0.f : 0.75f);
        float top[8], bot[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          top[e] = wx0 * va[e] + wx1 * vb[e];
          bot[e] = wx0 * vc[e] + wx1 * vd[e];
        }
        const int X0 = ox + halo;
        const int X1 = (halo > 0 && ox >= 1 && ox <= halo) ? halo - ox : -1;
        const int X2 = (halo > 0 && ox >= WO - 1 - halo && ox <= WO - 2) ? halo + 2 * (WO - 1) - ox : -1;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (!vk[k]) continue;
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = wyk[k][0] * top[e] + wyk[k][1] * bot[e];
          const uint4 ov = rp_pack8(o);
#pragma unroll
          for (int yi = 0; yi < 3; ++yi) {
            const int Y = Yk[k][yi];
            if (Y < 0) continue;
            bf16* rowp = ob + p.out.off_padded(n, Y, 0) + cv * 8;
            *reinterpret_cast<uint4*>(rowp + (long long)X0 * C) = ov;
            if (X1 >= 0) *reinterpret_cast<uint4*>(rowp + (long long)X1 * C) = ov;
            if (X2 >= 0) *reinterpret_cast<uint4*>(rowp + (long long)X2 * C) = ov;
          }
        }
      }
    }
  }
}

static bool row_up_ok(const dwc_hbuf_t* x, const dwc_hbuf_t* out) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DWC_ROWUP");
    on = e ? atoi(e) : 1;
  }
  const long long rowbytes = (long long)x->w * x->c * 2;
  return on && rowpipe_enabled() && x->dtype == DWC_BF16 && out->dtype == DWC_BF16 && x->layout == 0 && out->layout == 0 &&
         x->c % 8 == 0 && rowbytes >= 2048 && rowbytes <= 32768 && out->halo <= 2 * x->h - 1 && out->halo <= 2 * x->w - 1;
}

extern "C" int dwc_upsample_pad_fwd(const dwc_hbuf_t* x, const dwc_hbuf_t* out, dwc_stream_t stream) {
  DWC_CHECK(x->c % 8 == 0 && out->h == 2 * x->h && out->w == 2 * x->w && out->c == x->c && out->n == x->n,
            "dwc_upsample_pad_fwd: geometry mismatch");
  HB hx(*x), ho(*out);
  long long total = ho.padded_pixels() * (out->c / 8);
  if (row_up_ok(x, out)) {
    UpP up;
    up.x = hx; up.out = ho;
    up.rowbytes = x->w * x->c * 2;
    up.stages = 65536 / up.rowbytes;
    if (up.stages > 4) up.stages = 4;
    if (up.stages < 3) up.stages = 3;
    const size_t smem = (size_t)up.stages * up.rowbytes;
    static size_t attr = 0;
    if (smem > attr) {
      DWC_CUDA(cudaFuncSetAttribute(row_up_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
    int per_sm = 1;                                    // resident CTAs per SM (registers and shared memory): one wave
    DWC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, row_up_fwd_kernel, 256, smem));
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    int splits = (per_sm * dwc_num_sms()) / x->n;
    if (splits > x->h) splits = x->h;
    if (splits < 1) splits = 1;
    row_up_fwd_kernel<<<dim3(splits, x->n), 256, smem, as_stream(stream)>>>(up);
    DWC_LAUNCH_CHECK();
    return 0;
  }
  if (x->dtype == DWC_BF16 && ps_ok(out->c) && x->layout == 0 && out->layout == 0) {
    upsample_pad_fwd_fast_kernel<<<ps_grid(ho.hp * ho.wp, out->c / 8, out->n), 256, 0, as_stream(stream)>>>(hx, ho);
    DWC_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_T(x->dtype, (upsample_pad_fwd_kernel<T><<<ew_grid(total), 256, 0, as_stream(stream)>>>(hx, ho)));
  DWC_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwc_upsample_pad_bwd(const dwc_hbuf_t* dout, const dwc_hbuf_t* dx, int prefolded, dwc_stream_t stream) {
  DWC_CHECK(dx->c % 8 == 0 && dout->h == 2 * dx->h && dout->w == 2 * dx->w && dout->c == dx->c && dout->n == dx->n,
            "dwc_upsample_pad_bwd: geometry mismatch");
  HB hd(*dout), hx(*dx);
  if (prefolded) hd.refl = 0;
  if (dx->dtype == DWC_BF16 && hd.refl == 0 && ps_ok(dx->c) && dx->layout == 0 && dout->layout == 0) {
    upsample_pad_bwd_fast_kernel<<<ps_grid(hx.hp * hx.wp, dx->c / 8, dx->n), 256, 0, as_stream(stream)>>>(hd, hx);
    DWC_LAUNCH_CHECK();
    return 0;
  }
  long long total = hx.padded_pixels() * (dx->c / 8);
  DISPATCH_T(dx->dtype, (upsample_pad_bwd_kernel<T><<<ew_grid(total), 256, 0, as_stream(stream)>>>(hd, hx)));
  DWC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// NCHW fp32 image <-> haloed NHWC buffer (optional 2x2 average pooling)
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void image_pad_fwd_kernel(const float* __restrict__ img, int C, int H, int W, int pool, HB out) {
  const long long total = out.padded_pixels() * out.c;
  T* ob = reinterpret_cast<T*>(out.ptr);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % out.c);
    long long pix = i / out.c;
    int X = (int)(pix % out.wp);
    long long r = pix / out.wp;
    int Y = (int)(r % out.hp);
    int n = (int)(r / out.hp);
    int iy = reflect_idx(Y - out.halo, out.h), ix = reflect_idx(X - out.halo, out.w);
    float v = 0.f;
    if (c < C) {
      const float* p = img + ((long long)n * C + c) * H * W;
      if (pool == 1) v = p[(long long)iy * W + ix];
      else {
        // repeated F.interpolate(0.5, bilinear) == repeated 2x2 average == pool x pool average
        const float* q = p + (long long)(pool * iy) * W + pool * ix;
        float acc = 0.f;
        for (int a = 0; a < pool; ++a)
          for (int b = 0; b < pool; ++b) acc += q[(long long)a * W + b];
        v = acc / (float)(pool * pool);
      }
    }
    ob[out.off_padded(n, Y, X) + c] = from_f<T>(v);
  }
}

template <typename T>
__global__ void image_pad_bwd_kernel(HB dout, int pool, float* __restrict__ dimg, int C, int H, int W, int accumulate) {
  const long long total = (long long)dout.n * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % W);
    long long r = i / W;
    int y = (int)(r % H);
    r /= H;
    int c = (int)(r % C);
    int n = (int)(r / C);
    float g = fold_read1<T>(dout, n, y / pool, x / pool, c);
    if (pool > 1) g *= 1.f / (float)(pool * pool);
    dimg[i] = accumulate ? dimg[i] + g : g;
  }
}

extern "C" int dwc_image_pad_fwd(const float* img, int n, int c, int h, int w, int pool, const dwc_hbuf_t* out,
                                 dwc_stream_t stream) {
  DWC_CHECK(pool >= 1 && out->h * pool == h && out->w * pool == w && out->n == n && out->c >= c,
            "dwc_image_pad_fwd: geometry mismatch");
  HB ho(*out);
  long long total = ho.padded_pixels() * out->c;
  DISPATCH_T(out->dtype, (image_pad_fwd_kernel<T><<<ew_grid(total), 256, 0, as_stream(stream)>>>(img, c, h, w, pool, ho)));
  DWC_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwc_image_pad_bwd(const dwc_hbuf_t* dout, int pool, float* dimg, int n, int c, int h, int w,
                                 int accumulate, dwc_stream_t stream) {
  DWC_CHECK(pool >= 1 && dout->h * pool == h && dout->w * pool == w && dout->n == n && dout->c >= c,
            "dwc_image_pad_bwd: geometry mismatch");
  HB hd(*dout);
  long long total = (long long)n * c * h * w;
  DISPATCH_T(dout->dtype,
             (image_pad_bwd_kernel<T><<<ew_grid(total), 256, 0, as_stream(stream)>>>(hd, pool, dimg, c, h, w, accumulate)));
  DWC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// decoder heads (tanh | sigmoid) and attention blend
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void heads_fwd_kernel(HB y, float* __restrict__ img, float* __restrict__ att) {
  const long long hw = (long long)y.h * y.w;
  const long long total = (long long)y.n * hw;
  const T* yb = reinterpret_cast<const T*>(y.ptr);
  const int nc = y.c - 1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int n = (int)(i / hw);
    long long p = i - n * hw;
    int py = (int)(p / y.w), px = (int)(p % y.w);
    const T* q = yb + y.off(n, py, px);
    for (int c = 0; c < nc; ++c) img[((long long)n * nc + c) * hw + p] = tanhf(to_f<T>(q[c]));
    att[i] = 1.f / (1.f + __expf(-to_f<T>(q[nc])));
  }
}
template <typename T>
__global__ void heads_bwd_kernel(const float* __restrict__ dimg, const float* __restrict__ datt,
                                 const float* __restrict__ img, const float* __restrict__ att, HB dy) {
  const long long total = dy.padded_pixels();
  const long long hw = (long long)dy.h * dy.w;
  T* db = reinterpret_cast<T*>(dy.ptr);
  const int nc = dy.c - 1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int X = (int)(i % dy.wp);
    long long r = i / dy.wp;
    int Y = (int)(r % dy.hp);
    int n = (int)(r / dy.hp);
    int iy = Y - dy.halo, ix = X - dy.halo;
    T* q = db + dy.off_padded(n, Y, X);
    if (iy >= 0 && iy < dy.h && ix >= 0 && ix < dy.w) {
      long long p = (long long)iy * dy.w + ix;
      for (int c = 0; c < nc; ++c) {
        long long k = ((long long)n * nc + c) * hw + p;
        float t = img[k];
        q[c] = from_f<T>(dimg ? dimg[k] * (1.f - t * t) : 0.f);
      }
      float a = att[(long long)n * hw + p];
      q[nc] = from_f<T>(datt ? datt[(long long)n * hw + p] * a * (1.f - a) : 0.f);
    } else {
      for (int c = 0; c <= nc; ++c) q[c] = from_f<T>(0.f);
    }
  }
}
extern "C" int dwc_heads_fwd(const dwc_hbuf_t* y, float* img, float* att, dwc_stream_t stream) {
  HB hy(*y);
  long long total = (long long)y->n * y->h * y->w;
  DISPATCH_T(y->dtype, (heads_fwd_kernel<T><<<ew_grid(total), 256, 0, as_stream(stream)>>>(hy, img, att)));
  DWC_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwc_heads_bwd(const float* dimg, const float* datt, const float* img, const float* att,
                             const dwc_hbuf_t* dy, dwc_stream_t stream) {
  HB hd(*dy);
  DISPATCH_T(dy->dtype,
             (heads_bwd_kernel<T><<<ew_grid(hd.padded_pixels()), 256, 0, as_stream(stream)>>>(dimg, datt, img, att, hd)));
  DWC_LAUNCH_CHECK();
  return 0;
}

__global__ void blend_fwd_kernel(const float* __restrict__ img, const float* __restrict__ att,
                                 const float* __restrict__ real, float* __restrict__ out, int C, long long hw, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long n = i / (C * hw);
    long long p = i % hw;
    float a = att[n * hw + p];
    out[i] = img[i] * a + real[i] * (1.f - a);
  }
}
// one thread per (n, pixel): datt needs the sum over channels
__global__ void blend_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ img,
                                 const float* __restrict__ att, const float* __restrict__ real, float* __restrict__ dimg,
                                 float* __restrict__ datt, int C, long long hw, long long total_np) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_np; i += (long long)gridDim.x * blockDim.x) {
    long long n = i / hw, p = i % hw;
    float a = att[i], s = 0.f;
    for (int c = 0; c < C; ++c) {
      long long k = (n * C + c) * hw + p;
      float g = dout[k];
      dimg[k] = g * a;
      s += g * (img[k] - real[k]);
    }
    datt[i] = s;
  }
}
extern "C" int dwc_blend_fwd(const float* img, const float* att, const float* real, float* out, int n, int c, int hw,
                             dwc_stream_t stream) {
  long long total = (long long)n * c * hw;
  blend_fwd_kernel<<<ew_grid(total), 256, 0, as_stream(stream)>>>(img, att, real, out, c, hw, total);
  DWC_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwc_blend_bwd(const float* dout, const float* img, const float* att, const float* real, float* dimg,
                             float* datt, int n, int c, int hw, dwc_stream_t stream) {
  long long total = (long long)n * hw;
  blend_bwd_kernel<<<ew_grid(total), 256, 0, as_stream(stream)>>>(dout, img, att, real, dimg, datt, c, hw, total);
  DWC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// ReLU + global average pool (style-encoder tail)
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void relu_gap_fwd_kernel(HB y, float* __restrict__ out) {
  // one thread per (n, c)
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= y.n * y.c) return;
  int n = i / y.c, c = i % y.c;
  const T* yb = reinterpret_cast<const T*>(y.ptr);
  float s = 0.f;
  for (int py = 0; py < y.h; ++py)
    for (int px = 0; px < y.w; ++px) {
      float v = to_f<T>(yb[y.off(n, py, px) + c]);
      s += v > 0.f ? v : 0.f;
    }
  out[i] = s / (float)(y.h * y.w);
}
template <typename T>
__global__ void relu_gap_bwd_kernel(const float* __restrict__ dout, HB y, HB dy) {
  const long long total = dy.padded_pixels() * dy.c;
  const T* yb = reinterpret_cast<const T*>(y.ptr);
  T* db = reinterpret_cast<T*>(dy.ptr);
  const float inv = 1.f / (float)(y.h * y.w);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % dy.c);
    long long pix = i / dy.c;
    int X = (int)(pix % dy.wp);
    long long r = pix / dy.wp;
    int Y = (int)(r % dy.hp);
    int n = (int)(r / dy.hp);
    int iy = Y - dy.halo, ix = X - dy.halo;
    float g = 0.f;
    if (iy >= 0 && iy < dy.h && ix >= 0 && ix < dy.w) {
      float v = to_f<T>(yb[y.off(n, iy, ix) + c]);
      g = v > 0.f ? dout[(long long)n * dy.c + c] * inv : 0.f;
    }
    db[dy.off_padded(n, Y, X) + c] = from_f<T>(g);
  }
}
extern "C" int dwc_relu_gap_fwd(const dwc_hbuf_t* y, float* out, dwc_stream_t stream) {
  HB hy(*y);
  DISPATCH_T(y->dtype, (relu_gap_fwd_kernel<T><<<cdiv(y->n * y->c, 128), 128, 0, as_stream(stream)>>>(hy, out)));
  DWC_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwc_relu_gap_bwd(const float* dout, const dwc_hbuf_t* y, const dwc_hbuf_t* dy, dwc_stream_t stream) {
  HB hy(*y), hd(*dy);
  long long total = hd.padded_pixels() * dy->c;
  DISPATCH_T(y->dtype, (relu_gap_bwd_kernel<T><<<ew_grid(total), 256, 0, as_stream(stream)>>>(dout, hy, hd)));
  DWC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// row-im2col buffers that put the 3-channel / 4-channel convolutions on the tensor-core kernels
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void image_rows_fwd_kernel(const float* __restrict__ img, int N, int C, int H, int W, int pool, int pad, int sx,
                                      int ys, int Wo, T* __restrict__ rows) {
  const int h = H / pool, w = W / pool;               // interior after pooling
  const int Hp = h + 2 * pad, Wp = w + 2 * pad;
  const int Yr = Hp / ys;
  const long long total = (long long)N * ys * Yr * Wo * 8;      // one thread per (pixel window slot j)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int j = (int)(i & 7);
    long long r = i >> 3;
    int x = (int)(r % Wo); r /= Wo;
    int yr = (int)(r % Yr); r /= Yr;
    int z = (int)(r % ys);
    int n = (int)(r / ys);
    const int Y = yr * ys + z, X = x * sx + j;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (X < Wp) {
      const int iy = reflect_idx(Y - pad, h), ix = reflect_idx(X - pad, w);
      for (int ch = 0; ch < C && ch < 8; ++ch) {
        const float* p = img + ((long long)n * C + ch) * H * W;
        if (pool == 1) v[ch] = p[(long long)iy * W + ix];
        else {
          const float* q = p + (long long)(pool * iy) * W + pool * ix;
          float acc = 0.f;
          for (int a = 0; a < pool; ++a)
            for (int b = 0; b < pool; ++b) acc += q[(long long)a * W + b];
          v[ch] = acc / (float)(pool * pool);
        }
      }
    }
    Vec8<T>::store(rows + i * 8, v);
  }
}
extern "C" int dwc_image_rows_fwd(const float* img, int n, int c, int h, int w, int pool, int pad, int sx, int ys, int wo,
                                  void* rows, int dtype, dwc_stream_t stream) {
  DWC_CHECK(c <= 8 && pool >= 1 && (ys == 1 || ys == 2) && ((h / pool + 2 * pad) % ys) == 0, "dwc_image_rows_fwd: bad geometry");
  long long total = (long long)n * (h / pool + 2 * pad) * wo * 8;
  DISPATCH_T(dtype, (image_rows_fwd_kernel<T><<<ew_grid(total), 256, 0, as_stream(stream)>>>(
                        img, n, c, h, w, pool, pad, sx, ys, wo, reinterpret_cast<T*>(rows))));
  DWC_LAUNCH_CHECK();
  return 0;
}

// dy of the fused heads at interior pixel (n, y, x), 4 channels (0 outside the image)
__device__ __forceinline__ void heads_dy(const float* __restrict__ dimg, const float* __restrict__ datt,
                                         const float* __restrict__ img, const float* __restrict__ att, int n, int y, int x,
                                         int H, int W, float* d) {
  d[0] = d[1] = d[2] = d[3] = 0.f;
  if (y < 0 || y >= H || x < 0 || x >= W) return;
  const long long hw = (long long)H * W, p = (long long)y * W + x;
  if (dimg) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      long long k = ((long long)n * 3 + c) * hw + p;
      float t = img[k];
      d[c] = dimg[k] * (1.f - t * t);
    }
  }
  if (datt) {
    float a = att[(long long)n * hw + p];
    d[3] = datt[(long long)n * hw + p] * a * (1.f - a);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    heads_bwd_rows_kernel(const float* __restrict__ dimg, const float* __restrict__ datt, const float* __restrict__ img,
                          const float* __restrict__ att, int N, int H, int W, int halo, T* __restrict__ rows_d,
                          T* __restrict__ win, float* __restrict__ part) {
  const int Hh = H + 2 * halo, Wh = W + 2 * halo, Wu = W + halo;
  const long long n_rows = (long long)N * Hh * Wh * 8;      // rows_d slots
  const long long n_win = (long long)N * H * Wu * 8;        // win slots
  float bs[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_rows + n_win; i += (long long)gridDim.x * blockDim.x) {
    float v[8], d[4];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (i < n_rows) {
      int j = (int)(i & 7);
      long long r = i >> 3;
      int X = (int)(r % Wh); r /= Wh;
      int Y = (int)(r % Hh);
      int n = (int)(r / Hh);
      heads_dy(dimg, datt, img, att, n, Y - halo, X + j - halo, H, W, d);
      v[0] = d[0]; v[1] = d[1]; v[2] = d[2]; v[3] = d[3];
      Vec8<T>::store(rows_d + i * 8, v);
    } else {
      long long k = i - n_rows;
      int j = (int)(k & 7);
      long long r = k >> 3;
      int u = (int)(r % Wu); r /= Wu;
      int y = (int)(r % H);
      int n = (int)(r / H);
      heads_dy(dimg, datt, img, att, n, y, u - j, H, W, d);
      v[0] = d[0]; v[1] = d[1]; v[2] = d[2]; v[3] = d[3];
      Vec8<T>::store(win + k * 8, v);
      if (j == 0) { bs[0] += d[0]; bs[1] += d[1]; bs[2] += d[2]; bs[3] += d[3]; }   // each pixel once (u = x)
    }
  }
  __shared__ float red[4][256];
#pragma unroll
  for (int c = 0; c < 4; ++c) red[c][threadIdx.x] = bs[c];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
#pragma unroll
      for (int c = 0; c < 4; ++c) red[c][threadIdx.x] += red[c][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 4) part[(long long)blockIdx.x * 4 + threadIdx.x] = red[threadIdx.x][0];
}
extern "C" int dwc_heads_bwd_rows(const float* dimg, const float* datt, const float* img, const float* att, int n, int h,
                                  int w, int halo, void* rows_d, void* win, int dtype, float* part, int32_t* nblocks,
                                  dwc_stream_t stream) {
  long long total = (long long)n * (h + 2 * halo) * (w + 2 * halo) * 8 + (long long)n * h * (w + halo) * 8;
  int grid = ew_grid(total);
  if (grid > 1024) grid = 1024;
  *nblocks = grid;
  DISPATCH_T(dtype, (heads_bwd_rows_kernel<T><<<grid, 256, 0, as_stream(stream)>>>(
                        dimg, datt, img, att, n, h, w, halo, reinterpret_cast<T*>(rows_d), reinterpret_cast<T*>(win), part)));
  DWC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// per-(n,c) reduction entry points
// ---------------------------------------------------------------------------------------------------
extern "C" int dwc_nc_stats(const dwc_hbuf_t* y, int splits, float* stats, dwc_stream_t stream) {
  DWC_CHECK(y->c % 8 == 0 && y->layout == 0, "dwc_nc_stats: needs C %% 8 == 0 and plain layout");
  HB hy(*y);
  {
    RowP rp{};
    if (rowpipe_geom(y, &rp.nseg, &rp.segw, &rp.segbytes)) {
      rp.y = hy; rp.d = hy; rp.o1 = hy; rp.o2 = hy;
      rp.coef = nullptr; rp.bco = nullptr; rp.part = reinterpret_cast<float2*>(stats);
      rp.act = 0; rp.has_d = 0; rp.has_o2 = 0;
      return rowpipe_launch<RM_STATS>(rp, 1, splits, y->n, as_stream(stream));
    }
  }
  dim3 grid(cdiv(y->c, 32), splits, y->n);
  if (y->dtype == DWC_BF16 && ps_ok(y->c))
    nc_reduce_fast_kernel<0><<<dim3(splits, y->n), 256, 0, as_stream(stream)>>>(hy, hy, nullptr, 0, splits,
                                                                                reinterpret_cast<float2*>(stats));
  else if (y->dtype == DWC_BF16)
    nc_reduce_kernel<bf16, 0><<<grid, 256, 0, as_stream(stream)>>>(hy, hy, nullptr, 0, splits,
                                                                   reinterpret_cast<float2*>(stats));
  else
    nc_reduce_kernel<float, 0><<<grid, 256, 0, as_stream(stream)>>>(hy, hy, nullptr, 0, splits,
                                                                    reinterpret_cast<float2*>(stats));
  DWC_LAUNCH_CHECK();
  return 0;
}

// dwc_post_bwd_reduce(prefolded = 2) can fold dout's reflect-halo gradient itself (row-streaming kernel, whole padded
// rows of a plain-layout dout in one segment): no dwc_fold_halo launch in front of it, dout is folded afterwards.
// _can_fold: the geometry allows it.  _folds: ... and the product path uses it (DWC_FOLD_IN_REDUCE=1).  Measured on the
// training step: 32 launches fewer, but 19.91 vs 19.79 ms (+0.6 %): the step is bound by kernel throughput, not launch
// latency - the separate 7 us fold overlaps with other streams, the extra phase per row slows every reduction. Off.
extern "C" int dwc_post_bwd_reduce_can_fold(const dwc_hbuf_t* dout, const dwc_hbuf_t* y) {
  int nseg = 0, segw = 0, segbytes = 0;
  return dout->halo > 0 && dout->layout == 0 && dout->dtype == DWC_BF16 && y->dtype == DWC_BF16 &&
         dout->h >= 2 * dout->halo + 2 && dout->w >= 2 * dout->halo + 2 && dout->c % 8 == 0 &&
         rowpipe_geom(y, &nseg, &segw, &segbytes) && nseg == 1 &&
         (long long)(dout->w + 2 * dout->halo) * dout->c * 2 + segbytes <= 51000;
}
extern "C" int dwc_post_bwd_reduce_folds(const dwc_hbuf_t* dout, const dwc_hbuf_t* y) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DWC_FOLD_IN_REDUCE");
    on = e ? atoi(e) : 0;
  }
  return on && dwc_post_bwd_reduce_can_fold(dout, y);
}

extern "C" int dwc_post_bwd_reduce(const dwc_hbuf_t* dout, const dwc_hbuf_t* y, const float* coef, int act, int splits,
                                   float* red, int prefolded, dwc_stream_t stream) {
  DWC_CHECK(y->c % 8 == 0 && y->layout == 0, "dwc_post_bwd_reduce: needs C %% 8 == 0 and plain y");
  DWC_CHECK(dout->n == y->n && dout->h == y->h && dout->w == y->w && dout->c == y->c && dout->dtype == y->dtype,
            "dwc_post_bwd_reduce: geometry mismatch");
  HB hy(*y), hd(*dout);
  const bool fold_here = prefolded == 2;
  if (fold_here) {
    DWC_CHECK(dwc_post_bwd_reduce_can_fold(dout, y), "dwc_post_bwd_reduce: this site cannot fold the halo during the reduction");
    prefolded = 0;
  }
  if (prefolded) hd.refl = 0;
  {
    RowP rp{};
    if (fold_here && rowpipe_geom(y, &rp.nseg, &rp.segw, &rp.segbytes)) {
      rp.y = hy; rp.d = hd; rp.o1 = hy; rp.o2 = hy; rp.d_fold = 1;
      rp.coef = reinterpret_cast<const float4*>(coef); rp.bco = nullptr; rp.part = reinterpret_cast<float2*>(red);
      rp.act = act; rp.has_d = 1; rp.has_o2 = 0;
      return rowpipe_launch<RM_BRED>(rp, 2, splits, y->n, as_stream(stream));
    }
    if (rowpipe_geom(y, &rp.nseg, &rp.segw, &rp.segbytes) && (prefolded || dout->halo == 0) &&
        (dout->layout == 0 || rp.nseg == 1)) {
      rp.y = hy; rp.d = hd; rp.o1 = hy; rp.o2 = hy;
      rp.coef = reinterpret_cast<const float4*>(coef); rp.bco = nullptr; rp.part = reinterpret_cast<float2*>(red);
      rp.act = act; rp.has_d = 1; rp.has_o2 = 0;
      return rowpipe_launch<RM_BRED>(rp, 2, splits, y->n, as_stream(stream));
    }
  }
  dim3 grid(cdiv(y->c, 32), splits, y->n);
  if (y->dtype == DWC_BF16 && ps_ok(y->c))
    nc_reduce_fast_kernel<1><<<dim3(splits, y->n), 256, 0, as_stream(stream)>>>(
        hy, hd, reinterpret_cast<const float4*>(coef), act, splits, reinterpret_cast<float2*>(red));
  else if (y->dtype == DWC_BF16)
    nc_reduce_kernel<bf16, 1><<<grid, 256, 0, as_stream(stream)>>>(hy, hd, reinterpret_cast<const float4*>(coef), act,
                                                                   splits, reinterpret_cast<float2*>(red));
  else
    nc_reduce_kernel<float, 1><<<grid, 256, 0, as_stream(stream)>>>(hy, hd, reinterpret_cast<const float4*>(coef), act,
                                                                    splits, reinterpret_cast<float2*>(red));
  DWC_LAUNCH_CHECK();
  return 0;
}
