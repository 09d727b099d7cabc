// GMM style sampling / KL, scalar losses, Adam + EMA, weight packing, casts.
#include "common.cuh"

static inline int grid1d(long long count) {
  long long b = (count + 255) / 256;
  return (int)(b > 2368 ? 2368 : (b < 1 ? 1 : b));
}

__device__ __forceinline__ float block_sum_f(float v) {
  __shared__ float sm[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < (blockDim.x >> 5) ? sm[lane] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;   // valid in warp 0
}

// ---------------------------------------------------------------------------------------------------
// GMM
// ---------------------------------------------------------------------------------------------------
__global__ void gmm_sample_kernel(const float* __restrict__ mu, const float* __restrict__ eps, float stddev,
                                  float* __restrict__ z, int B, int ncls, int cdim) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * ncls * cdim) return;
  int k = i % cdim;
  int j = (i / cdim) % ncls;
  int b = i / (cdim * ncls);
  // two roundings (product, then sum) like torch.normal(mean, std) = randn().mul_(std).add_(mean): bit-identical to
  // tdist.Normal(mu, stddev).sample() of the reference for any stddev, not only powers of two (no FMA contraction)
  z[i] = __fadd_rn(mu[b * ncls + j], __fmul_rn(stddev, eps[((long long)k * B + b) * ncls + j]));
}
extern "C" int dwc_gmm_sample(const float* mu, const float* eps, float stddev, float* z, int b, int ncls, int cdim,
                              dwc_stream_t stream) {
  gmm_sample_kernel<<<cdiv(b * ncls * cdim, 128), 128, 0, as_stream(stream)>>>(mu, eps, stddev, z, b, ncls, cdim);
  DWC_LAUNCH_CHECK();
  return 0;
}

// one warp per sample row; loss = (1/B) sum_b sum_{j,k} 0.5*(log(sigma) - lv + (exp(lv) + (mu-c)^2)/sigma - 1)
__global__ void gmm_kl_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ c,
                              float sigma, float* __restrict__ loss, float* __restrict__ dmu, float* __restrict__ dlv,
                              int B, int ncls, int cdim) {
  // single block: deterministic
  const int D = ncls * cdim;
  const float lsig = logf(sigma), inv = 1.f / sigma, invB = 1.f / (float)B;
  float s = 0.f;
  for (int i = threadIdx.x; i < B * D; i += blockDim.x) {
    int b = i / D, j = (i % D) / cdim;
    float m = mu[i], l = lv[i], d = m - c[b * ncls + j];
    float v = expf(l);
    s += 0.5f * (lsig - l + (v + d * d) * inv - 1.f);
    dmu[i] = d * inv * invB;
    dlv[i] = 0.5f * (v * inv - 1.f) * invB;
  }
  s = block_sum_f(s);
  if (threadIdx.x == 0) loss[0] = s * invB;
}
extern "C" int dwc_gmm_kl(const float* mu, const float* lv, const float* c, float sigma, float* loss, float* dmu,
                          float* dlv, int b, int ncls, int cdim, dwc_stream_t stream) {
  gmm_kl_kernel<<<1, 256, 0, as_stream(stream)>>>(mu, lv, c, sigma, loss, dmu, dlv, b, ncls, cdim);
  DWC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// scalar losses.  The L1 forward is a two-level sum with a fixed order (bit-reproducible): every block parks its
// partial sum in the caller's zeroed scratch, the block that draws the last ticket adds them by index.
// loss[0] result, loss[1] ticket, loss[DWC_L1_SCRATCH_OFF + block] partial sums.
// ---------------------------------------------------------------------------------------------------
constexpr int L1_MAX_BLOCKS = 592;
constexpr int L1_PART_OFF = 8;
template <typename TA, typename TB>
__global__ void l1_fwd_kernel(const TA* __restrict__ a, const TB* __restrict__ b, long long count, float inv,
                              float* __restrict__ loss) {
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    s += fabsf(to_f<TA>(a[i]) - to_f<TB>(b[i]));
  s = block_sum_f(s);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __stcg(loss + L1_PART_OFF + blockIdx.x, s);
    __threadfence();
    last = atomicAdd(reinterpret_cast<unsigned int*>(loss + 1), 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(loss + L1_PART_OFF + i);
  t = block_sum_f(t);
  if (threadIdx.x == 0) loss[0] = t * inv;
}
template <typename TA, typename TB>
__global__ void l1_bwd_kernel(const TA* __restrict__ a, const TB* __restrict__ b, long long count, float inv,
                              const float* __restrict__ gscale, TA* __restrict__ da, TB* __restrict__ db) {
  const float g = gscale[0] * inv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    float d = to_f<TA>(a[i]) - to_f<TB>(b[i]);
    float s = d > 0.f ? g : (d < 0.f ? -g : 0.f);
    if (da) da[i] = from_f<TA>(s);
    if (db) db[i] = from_f<TB>(-s);
  }
}
extern "C" int dwc_l1_loss_fwd(const void* a, int a_dtype, const void* b, int b_dtype, int64_t count, float* loss,
                               dwc_stream_t stream) {
  const float inv = 1.f / (float)count;
  int g = grid1d(count) > L1_MAX_BLOCKS ? L1_MAX_BLOCKS : grid1d(count);
  cudaStream_t st = as_stream(stream);
  if (a_dtype == DWC_F32 && b_dtype == DWC_F32)
    l1_fwd_kernel<float, float><<<g, 256, 0, st>>>((const float*)a, (const float*)b, count, inv, loss);
  else if (a_dtype == DWC_BF16 && b_dtype == DWC_BF16)
    l1_fwd_kernel<bf16, bf16><<<g, 256, 0, st>>>((const bf16*)a, (const bf16*)b, count, inv, loss);
  else if (a_dtype == DWC_BF16)
    l1_fwd_kernel<bf16, float><<<g, 256, 0, st>>>((const bf16*)a, (const float*)b, count, inv, loss);
  else
    l1_fwd_kernel<float, bf16><<<g, 256, 0, st>>>((const float*)a, (const bf16*)b, count, inv, loss);
  DWC_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwc_l1_loss_bwd(const void* a, int a_dtype, const void* b, int b_dtype, int64_t count,
                               const float* gscale, void* da, void* db, dwc_stream_t stream) {
  const float inv = 1.f / (float)count;
  int g = grid1d(count);
  cudaStream_t st = as_stream(stream);
  if (a_dtype == DWC_F32 && b_dtype == DWC_F32)
    l1_bwd_kernel<float, float><<<g, 256, 0, st>>>((const float*)a, (const float*)b, count, inv, gscale, (float*)da, (float*)db);
  else if (a_dtype == DWC_BF16 && b_dtype == DWC_BF16)
    l1_bwd_kernel<bf16, bf16><<<g, 256, 0, st>>>((const bf16*)a, (const bf16*)b, count, inv, gscale, (bf16*)da, (bf16*)db);
  else if (a_dtype == DWC_BF16)
    l1_bwd_kernel<bf16, float><<<g, 256, 0, st>>>((const bf16*)a, (const float*)b, count, inv, gscale, (bf16*)da, (float*)db);
  else
    l1_bwd_kernel<float, bf16><<<g, 256, 0, st>>>((const float*)a, (const bf16*)b, count, inv, gscale, (float*)da, (bf16*)db);
  DWC_LAUNCH_CHECK();
  return 0;
}

// mode 0: mean (x-target)^2 ; mode 1: mean BCE-with-logits(x, y)
__global__ void small_loss_fwd_kernel(int mode, const float* __restrict__ x, const float* __restrict__ y, float target,
                                      long long count, float* __restrict__ loss) {
  float s = 0.f;
  for (long long i = threadIdx.x; i < count; i += blockDim.x) {
    float v = x[i];
    if (mode == 0) s += (v - target) * (v - target);
    else s += fmaxf(v, 0.f) - v * y[i] + log1pf(expf(-fabsf(v)));
  }
  s = block_sum_f(s);
  if (threadIdx.x == 0) loss[0] = s / (float)count;
}
__global__ void small_loss_bwd_kernel(int mode, const float* __restrict__ x, const float* __restrict__ y, float target,
                                      long long count, const float* __restrict__ gscale, float* __restrict__ dx) {
  const float g = gscale[0] / (float)count;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    dx[i] = mode == 0 ? 2.f * (v - target) * g : (1.f / (1.f + expf(-v)) - y[i]) * g;
  }
}
extern "C" int dwc_mse_const_loss_fwd(const float* x, float target, int64_t count, float* loss, dwc_stream_t stream) {
  small_loss_fwd_kernel<<<1, 256, 0, as_stream(stream)>>>(0, x, nullptr, target, count, loss);
  DWC_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwc_mse_const_loss_bwd(const float* x, float target, int64_t count, const float* gscale, float* dx,
                                      dwc_stream_t stream) {
  small_loss_bwd_kernel<<<grid1d(count), 256, 0, as_stream(stream)>>>(0, x, nullptr, target, count, gscale, dx);
  DWC_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwc_bce_logits_loss_fwd(const float* x, const float* y, int64_t count, float* loss, dwc_stream_t stream) {
  small_loss_fwd_kernel<<<1, 256, 0, as_stream(stream)>>>(1, x, y, 0.f, count, loss);
  DWC_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwc_bce_logits_loss_bwd(const float* x, const float* y, int64_t count, const float* gscale, float* dx,
                                       dwc_stream_t stream) {
  small_loss_bwd_kernel<<<grid1d(count), 256, 0, as_stream(stream)>>>(1, x, y, 0.f, count, gscale, dx);
  DWC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// One discriminator scale's adversarial terms in one kernel per direction: up to 8 weighted terms, each the mean over
// a row range of either (src - target)^2 (LSGAN) or BCE-with-logits(cls, labels) - instead of a slice copy, a loss
// kernel, a zero-fill, a slice scatter and a gradient add per term (networks.py:116-170, solver.py:206-207,333-334).
// ---------------------------------------------------------------------------------------------------
struct AdvTerms {
  int n;
  dwc_adv_term_t t[DWC_ADV_MAX_TERMS];
};
__global__ void __launch_bounds__(256) adv_loss_fwd_kernel(const float* __restrict__ src, int src_cols,
                                                           const float* __restrict__ cls, int cls_cols,
                                                           const float* __restrict__ labels, const AdvTerms T,
                                                           float* __restrict__ loss) {
  float total = 0.f;
  for (int k = 0; k < T.n; ++k) {                      // single block, fixed order: deterministic
    const dwc_adv_term_t t = T.t[k];
    const int cols = t.kind == 0 ? src_cols : cls_cols;
    const long long count = (long long)(t.row1 - t.row0) * cols;
    const float* x = (t.kind == 0 ? src : cls) + (long long)t.row0 * cols;
    float s = 0.f;
    for (long long i = threadIdx.x; i < count; i += blockDim.x) {
      const float v = x[i];
      if (t.kind == 0) s += (v - t.target) * (v - t.target);
      else s += fmaxf(v, 0.f) - v * labels[i] + log1pf(expf(-fabsf(v)));
    }
    s = block_sum_f(s);
    total += t.weight * (s / (float)count);
  }
  if (threadIdx.x == 0) loss[0] = total;
}
__global__ void __launch_bounds__(256) adv_loss_bwd_kernel(const float* __restrict__ src, int src_rows, int src_cols,
                                                           const float* __restrict__ cls, int cls_rows, int cls_cols,
                                                           const float* __restrict__ labels, const AdvTerms T,
                                                           const float* __restrict__ gscale, float* __restrict__ dsrc,
                                                           float* __restrict__ dcls) {
  const float g = gscale[0];
  const long long ns = (long long)src_rows * src_cols, nc = (long long)cls_rows * cls_cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ns + nc; i += (long long)gridDim.x * blockDim.x) {
    const bool is_src = i < ns;
    const long long j = is_src ? i : i - ns;
    const int cols = is_src ? src_cols : cls_cols;
    const int row = (int)(j / cols);
    const float v = is_src ? src[j] : cls[j];
    float d = 0.f;
    for (int k = 0; k < T.n; ++k) {
      const dwc_adv_term_t t = T.t[k];
      if ((t.kind == 0) != is_src || row < t.row0 || row >= t.row1) continue;
      const float gk = g * t.weight / (float)((long long)(t.row1 - t.row0) * cols);
      d += t.kind == 0 ? 2.f * (v - t.target) * gk
                       : (1.f / (1.f + expf(-v)) - labels[j - (long long)t.row0 * cols]) * gk;
    }
    if (is_src) dsrc[j] = d;
    else dcls[j] = d;
  }
}
static int adv_terms(const dwc_adv_term_t* terms, int nterms, int src_rows, int cls_rows, AdvTerms* T) {
  DWC_CHECK(terms && nterms > 0 && nterms <= DWC_ADV_MAX_TERMS, "dwc_adv_loss: 1..%d terms", DWC_ADV_MAX_TERMS);
  T->n = nterms;
  for (int k = 0; k < nterms; ++k) {
    T->t[k] = terms[k];
    const int rows = terms[k].kind == 0 ? src_rows : cls_rows;
    DWC_CHECK((terms[k].kind == 0 || terms[k].kind == 1) && terms[k].row0 >= 0 && terms[k].row1 > terms[k].row0 &&
                  terms[k].row1 <= rows, "dwc_adv_loss: bad term %d", k);
  }
  return 0;
}
extern "C" int dwc_adv_loss_fwd(const float* src, int src_rows, int src_cols, const float* cls, int cls_rows, int cls_cols,
                                const float* labels, const dwc_adv_term_t* terms, int nterms, float* loss,
                                dwc_stream_t stream) {
  AdvTerms T;
  if (adv_terms(terms, nterms, src_rows, cls_rows, &T)) return 1;
  adv_loss_fwd_kernel<<<1, 256, 0, as_stream(stream)>>>(src, src_cols, cls, cls_cols, labels, T, loss);
  DWC_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwc_adv_loss_bwd(const float* src, int src_rows, int src_cols, const float* cls, int cls_rows, int cls_cols,
                                const float* labels, const dwc_adv_term_t* terms, int nterms, const float* gscale,
                                float* dsrc, float* dcls, dwc_stream_t stream) {
  AdvTerms T;
  if (adv_terms(terms, nterms, src_rows, cls_rows, &T)) return 1;
  const long long total = (long long)src_rows * src_cols + (long long)cls_rows * cls_cols;
  adv_loss_bwd_kernel<<<grid1d(total), 256, 0, as_stream(stream)>>>(src, src_rows, src_cols, cls, cls_rows, cls_cols, labels,
                                                                     T, gscale, dsrc, dcls);
  DWC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Adam (coupled L2) and EMA over flat fp32 buffers
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                long long count, const uint8_t* __restrict__ active, const float* __restrict__ hyper) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], bc1 = hyper[5],
              bc2 = hyper[6], gs = hyper[7];
  const float step = lr / bc1, rs = rsqrtf(bc2);
  // one block per 1024-element chunk
  const long long base = (long long)blockIdx.x * 1024;
  if (active && !active[blockIdx.x]) return;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    long long i = base + threadIdx.x + k * 256;
    if (i < count) {
      float pi = p[i];
      float gi = g[i] * gs + wd * pi;
      float mi = b1 * m[i] + (1.f - b1) * gi;
      float vi = b2 * v[i] + (1.f - b2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      p[i] = pi - step * mi / (sqrtf(vi) * rs + eps);
    }
  }
}
extern "C" int dwc_adam_step(float* param, const float* grad, float* m, float* v, int64_t count, const uint8_t* active,
                             const float* hyper, dwc_stream_t stream) {
  adam_kernel<<<cdiv(count, 1024), 256, 0, as_stream(stream)>>>(param, grad, m, v, count, active, hyper);
  DWC_LAUNCH_CHECK();
  return 0;
}
__global__ void ema_kernel(const float* __restrict__ p, float* __restrict__ avg, long long count, float beta) {
  // torch.lerp(p, avg, beta) with beta >= 0.5: avg - (avg - p) * (1 - beta)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    float a = avg[i], q = p[i];
    avg[i] = beta >= 0.5f ? a - (a - q) * (1.f - beta) : q + beta * (a - q);
  }
}
extern "C" int dwc_ema_step(const float* param, float* avg, int64_t count, float beta, dwc_stream_t stream) {
  ema_kernel<<<grid1d(count), 256, 0, as_stream(stream)>>>(param, avg, count, beta);
  DWC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// weight packing (master fp32 [Cout][KH][KW][Cin] -> GEMM operands)
// ---------------------------------------------------------------------------------------------------
// value of element i of the packed operand
__device__ __forceinline__ float pack_elem(const float* __restrict__ w, int Cout, int KH, int KW, int Cin, int mode,
                                           int rows_padded, long long i) {
  const int taps = KH * KW;
  float v = 0.f;
  if (mode == 0) {
    // out[row=co][t][ci]
    long long K = (long long)taps * Cin;
    int row = (int)(i / K);
    if (row < Cout) v = w[i];
  } else if (mode == 1) {
    // out[row=ci][t'][co] = w[co][taps-1-t'][ci]
    long long K = (long long)taps * Cout;
    int row = (int)(i / K);
    long long r = i % K;
    int tp = (int)(r / Cout), co = (int)(r % Cout);
    if (row < Cin) v = w[((long long)co * taps + (taps - 1 - tp)) * Cin + row];
  } else if (mode == 3) {
    // out[row=co][kh][j*8+ci] = w[co][kh][j][ci]
    long long K = (long long)KH * 64;
    int row = (int)(i / K);
    int r = (int)(i % K);
    int kh = r >> 6, j = (r >> 3) & 7, ci = r & 7;
    if (row < Cout && j < KW && ci < Cin) v = w[(((long long)row * KH + kh) * KW + j) * Cin + ci];
  } else if (mode == 4) {
    // out[row=ci][kh'][j*8+co] = w[co][KH-1-kh'][KW-1-j][ci]
    long long K = (long long)KH * 64;
    int row = (int)(i / K);
    int r = (int)(i % K);
    int khp = r >> 6, j = (r >> 3) & 7, co = r & 7;
    if (row < Cin && j < KW && co < Cout) v = w[(((long long)co * KH + (KH - 1 - khp)) * KW + (KW - 1 - j)) * Cin + row];
  } else {
    // 4 phases: out[phase][row=ci][(i',j')][co] = w[co][2(1-i')+py][2(1-j')+px][ci]   (KH = KW = 4)
    long long K = 4LL * Cout;
    long long per_phase = (long long)rows_padded * K;
    int phase = (int)(i / per_phase);
    long long r = i % per_phase;
    int row = (int)(r / K);
    r %= K;
    int tp = (int)(r / Cout), co = (int)(r % Cout);
    int ip = tp >> 1, jp = tp & 1, py = phase >> 1, px = phase & 1;
    int kh = 2 * (1 - ip) + py, kw = 2 * (1 - jp) + px;
    if (row < Cin) v = w[(((long long)co * KH + kh) * KW + kw) * Cin + row];
  }
  return v;
}

template <typename T>
__global__ void pack_weights_kernel(const float* __restrict__ w, int Cout, int KH, int KW, int Cin, int mode,
                                    T* __restrict__ out, int rows_padded, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    out[i] = from_f<T>(pack_elem(w, Cout, KH, KW, Cin, mode, rows_padded, i));
}

// every packed operand of a network in ONE launch: blockIdx.y = table entry (the table lives in device memory), the
// blocks of a row stride over the entry's work units.  bf16 outputs take coalesced paths: mode 0 is a cast (8 elements
// per thread, 16-byte stores); the data-gradient operands (modes 1 / 2) are per-tap [Cout x Cin] -> [Cin x Cout]
// transposes done through a 32 x 33 shared-memory tile, so that both the fp32 reads (along Cin) and the bf16 writes
// (along Cout) are contiguous; the small row-im2col modes and fp32 outputs keep the element-wise path.
__global__ void __launch_bounds__(256) pack_weights_batch_kernel(const dwc_pack_entry_t* __restrict__ table) {
  __shared__ float tile[32][33];
  const dwc_pack_entry_t e = table[blockIdx.y];
  const float* __restrict__ w = e.w;
  const int taps = e.kh * e.kw;
  if (e.out_dtype == DWC_BF16 && e.mode == 0 && (e.total & 7) == 0 && ((long long)taps * e.cin) % 8 == 0) {
    bf16* out = reinterpret_cast<bf16*>(e.out);
    const long long valid = (long long)e.cout * taps * e.cin;          // rows beyond cout are zero padding
    for (long long i = (blockIdx.x * 256LL + threadIdx.x) * 8; i < e.total; i += gridDim.x * 2048LL) {
      float f[8];
      if (i < valid) {
        const float4 a = *reinterpret_cast<const float4*>(w + i), b = *reinterpret_cast<const float4*>(w + i + 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
      Vec8<bf16>::store(out + i, f);
    }
    return;
  }
  if (e.out_dtype == DWC_BF16 && (e.mode == 1 || (e.mode == 2 && e.kh == 4 && e.kw == 4)) && (e.cout & 1) == 0) {
    bf16* out = reinterpret_cast<bf16*>(e.out);
    const int rows = e.rows_padded, Cout = e.cout, Cin = e.cin;
    const int ct = (rows + 31) / 32, ot = (Cout + 31) / 32;              // tiles along ci (incl. padding rows) and co
    const int units = taps * ct * ot;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;              // 32 x 8 threads
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int t = u / (ct * ot), r = u - t * (ct * ot);
      const int ci0 = (r / ot) * 32, co0 = (r % ot) * 32;
      // read w[co][t][ci]: ci contiguous
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = co0 + ty + 8 * j, ci = ci0 + tx;
        tile[ty + 8 * j][tx] = (co < Cout && ci < Cin) ? w[((long long)co * taps + t) * Cin + ci] : 0.f;
      }
      __syncthreads();
      // where tap t of the forward filter lands in the data-gradient operand
      long long base;
      long long K;
      if (e.mode == 1) {
        K = (long long)taps * Cout;
        base = (long long)(taps - 1 - t) * Cout;                         // flipped tap
      } else {
        const int kh = t >> 2, kw = t & 3;
        const int py = kh & 1, ip = 1 - (kh >> 1), px = kw & 1, jp = 1 - (kw >> 1);
        K = 4LL * Cout;
        base = (long long)(py * 2 + px) * rows * K + (long long)(ip * 2 + jp) * Cout;
      }
      // write out[ci][...][co]: co contiguous, two bf16 per thread
      const int cw = (threadIdx.x & 15) * 2, rw = threadIdx.x >> 4;       // 16 column pairs x 16 rows per pass
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int ci = ci0 + rw + 16 * j, co = co0 + cw;
        if (ci < rows && co < Cout)
          *reinterpret_cast<__nv_bfloat162*>(out + base + (long long)ci * K + co) =
              __floats2bfloat162_rn(tile[cw][rw + 16 * j], tile[cw + 1][rw + 16 * j]);
      }
      __syncthreads();
    }
    return;
  }
  if (e.out_dtype == DWC_F32) {
    float* out = reinterpret_cast<float*>(e.out);
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < e.total; i += gridDim.x * 256LL)
      out[i] = pack_elem(w, e.cout, e.kh, e.kw, e.cin, e.mode, e.rows_padded, i);
  } else {
    bf16* out = reinterpret_cast<bf16*>(e.out);
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < e.total; i += gridDim.x * 256LL)
      out[i] = __float2bfloat16_rn(pack_elem(w, e.cout, e.kh, e.kw, e.cin, e.mode, e.rows_padded, i));
  }
}
extern "C" int dwc_pack_weights_batch(const dwc_pack_entry_t* table_dev, int count, dwc_stream_t stream) {
  DWC_CHECK(table_dev != nullptr && count > 0 && count <= 65535, "dwc_pack_weights_batch: bad table");
  pack_weights_batch_kernel<<<dim3(128, count), 256, 0, as_stream(stream)>>>(table_dev);
  DWC_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwc_pack_weights(const float* w, int cout, int taps_h, int taps_w, int cin, int mode, void* out,
                                int out_dtype, int rows_padded, dwc_stream_t stream) {
  DWC_CHECK(mode >= 0 && mode <= 4, "dwc_pack_weights: bad mode");
  DWC_CHECK(mode < 3 || (taps_w <= 8 && (mode == 3 ? cin : cout) <= 8), "dwc_pack_weights: row-im2col modes need kw, c <= 8");
  DWC_CHECK(mode != 2 || (taps_h == 4 && taps_w == 4), "dwc_pack_weights: mode 2 needs a 4x4 kernel");
  long long total;
  if (mode == 0) total = (long long)rows_padded * taps_h * taps_w * cin;
  else if (mode == 1) total = (long long)rows_padded * taps_h * taps_w * cout;
  else if (mode == 3 || mode == 4) total = (long long)rows_padded * taps_h * 64;
  else total = 4LL * rows_padded * 4 * cout;
  if (out_dtype == DWC_F32)
    pack_weights_kernel<float><<<grid1d(total), 256, 0, as_stream(stream)>>>(w, cout, taps_h, taps_w, cin, mode,
                                                                              (float*)out, rows_padded, total);
  else
    pack_weights_kernel<bf16><<<grid1d(total), 256, 0, as_stream(stream)>>>(w, cout, taps_h, taps_w, cin, mode,
                                                                             (bf16*)out, rows_padded, total);
  DWC_LAUNCH_CHECK();
  return 0;
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, long long count) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    d[i] = from_f<TD>(to_f<TS>(s[i]));
}
extern "C" int dwc_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t count, dwc_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  int g = grid1d(count);
  if (src_dtype == DWC_F32 && dst_dtype == DWC_BF16) cast_kernel<float, bf16><<<g, 256, 0, st>>>((const float*)src, (bf16*)dst, count);
  else if (src_dtype == DWC_BF16 && dst_dtype == DWC_F32) cast_kernel<bf16, float><<<g, 256, 0, st>>>((const bf16*)src, (float*)dst, count);
  else if (src_dtype == DWC_F32) cast_kernel<float, float><<<g, 256, 0, st>>>((const float*)src, (float*)dst, count);
  else cast_kernel<bf16, bf16><<<g, 256, 0, st>>>((const bf16*)src, (bf16*)dst, count);
  DWC_LAUNCH_CHECK();
  return 0;
}
template <typename T> __global__ void fill_kernel(T* d, float v, long long count) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    d[i] = from_f<T>(v);
}
extern "C" int dwc_fill(void* dst, int dtype, float value, int64_t count, dwc_stream_t stream) {
  if (dtype == DWC_F32) fill_kernel<float><<<grid1d(count), 256, 0, as_stream(stream)>>>((float*)dst, value, count);
  else fill_kernel<bf16><<<grid1d(count), 256, 0, as_stream(stream)>>>((bf16*)dst, value, count);
  DWC_LAUNCH_CHECK();
  return 0;
}
