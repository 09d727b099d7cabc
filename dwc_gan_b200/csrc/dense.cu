// Small dense algebra and the text encoder: strided fp32 GEMM (style MLP, mapping network, heads,
// LSTM input projections and weight gradients), embedding + style concat (the LSTM recurrence is in lstm.cu).
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------
// C[m,n] = act(alpha * sum_k A[m,k] B[k,n] + bias[n] + beta*C[m,n]); 64x64 tile, 256 threads x (4x4)
// ---------------------------------------------------------------------------------------------------
// blockIdx.z = K split: split z covers k in [z*kper, min(K, (z+1)*kper)) and (when gridDim.z > 1) writes its raw
// partial sums to a dense [split][M][N] workspace; sgemm_splitk_finish_kernel applies alpha / bias / beta / act.
template <typename TA>
__global__ void __launch_bounds__(256)
    sgemm_kernel(int M, int N, int K, float alpha, const TA* __restrict__ A, long long a_sm, long long a_sk,
                 const float* __restrict__ B, long long b_sk, long long b_sn, float beta, float* __restrict__ C,
                 long long c_sm, long long c_sn, const float* __restrict__ bias, int act, int kper,
                 float* __restrict__ ws) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loaders: choose the thread->element map so that the unit-stride axis is the fast one
  const bool a_kfast = (a_sk == 1);
  const bool b_nfast = (b_sn == 1);
  const int kbeg = blockIdx.z * kper, kend = min(K, kbeg + kper);
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int e = tid + j * 256;   // 0..1023
      int mm, kk;
      if (a_kfast) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < kend) ? to_f<TA>(A[gm * a_sm + gk * a_sk]) : 0.f;
      int nn, k2;
      if (b_nfast) { nn = e & 63; k2 = e >> 6; } else { k2 = e & 15; nn = e >> 4; }
      int gn = n0 + nn, gk2 = k0 + k2;
      Bs[k2][nn] = (gn < N && gk2 < kend) ? B[gk2 * b_sk + gn * b_sn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      if (split) {
        ws[((long long)blockIdx.z * M + m) * N + n] = acc[i][j];
      } else {
        float* o = C + m * c_sm + n * c_sn;
        float v = alpha * acc[i][j] + (bias ? bias[n] : 0.f) + (beta != 0.f ? beta * *o : 0.f);
        if (act == 1) v = v > 0.f ? v : 0.f;
        *o = v;
      }
    }
  }
}

// Large problems (LSTM input projections and their gradients): 128 x 128 tile, 256 threads x (8 x 8), K in chunks of 16
// with register prefetch of the next chunk (double-buffered shared memory).
template <typename TA>
__global__ void __launch_bounds__(256)
    sgemm128_kernel(int M, int N, int K, float alpha, const TA* __restrict__ A, long long a_sm, long long a_sk,
                    const float* __restrict__ B, long long b_sk, long long b_sn, float beta, float* __restrict__ C,
                    long long c_sm, long long c_sn, const float* __restrict__ bias, int act) {
  __shared__ float As[2][16][128 + 4];
  __shared__ float Bs[2][16][128 + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * 128;
  const int ty = tid >> 4, tx = tid & 15;          // rows {ty*4.., 64+ty*4..}, cols {tx*4.., 64+tx*4..}
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (a_sk == 1);
  const bool b_nfast = (b_sn == 1);
  float ra[8], rb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = tid + j * 256;   // 0..2047
      int mm, kk;
      if (a_kfast) { kk = e & 15; mm = e >> 4; } else { mm = e & 127; kk = e >> 7; }
      const int gm = m0 + mm, gk = k0 + kk;
      ra[j] = (gm < M && gk < K) ? to_f<TA>(A[gm * a_sm + gk * a_sk]) : 0.f;
      int nn, k2;
      if (b_nfast) { nn = e & 127; k2 = e >> 7; } else { k2 = e & 15; nn = e >> 4; }
      const int gn = n0 + nn, gk2 = k0 + k2;
      rb[j] = (gn < N && gk2 < K) ? B[gk2 * b_sk + gn * b_sn] : 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = tid + j * 256;
      int mm, kk;
      if (a_kfast) { kk = e & 15; mm = e >> 4; } else { mm = e & 127; kk = e >> 7; }
      As[buf][kk][mm] = ra[j];
      int nn, k2;
      if (b_nfast) { nn = e & 127; k2 = e >> 7; } else { k2 = e & 15; nn = e >> 4; }
      Bs[buf][k2][nn] = rb[j];
    }
  };
  fetch(0);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += 16) {
    const bool more = k0 + 16 < K;
    if (more) fetch(k0 + 16);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      stash(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (n >= N) continue;
      float* o = C + m * c_sm + n * c_sn;
      float v = alpha * acc[i][j] + (bias ? bias[n] : 0.f) + (beta != 0.f ? beta * *o : 0.f);
      if (act == 1) v = v > 0.f ? v : 0.f;
      *o = v;
    }
  }
}

__global__ void sgemm_splitk_finish_kernel(int M, int N, int splits, float alpha, const float* __restrict__ ws,
                                           float beta, float* __restrict__ C, long long c_sm, long long c_sn,
                                           const float* __restrict__ bias, int act) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const int m = i / N, n = i - m * N;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += ws[(long long)z * M * N + i];   // fixed order: deterministic
  float* o = C + m * c_sm + n * c_sn;
  float v = alpha * s + (bias ? bias[n] : 0.f) + (beta != 0.f ? beta * *o : 0.f);
  if (act == 1) v = v > 0.f ? v : 0.f;
  *o = v;
}

// number of K splits for a problem whose (M, N) tile grid alone cannot fill the GPU
static int sgemm_splits(int m, int n, int k) {
  // The 64 x 64 kernel has one K chunk in flight per CTA, i.e. it is latency-bound unless several CTAs share an SM:
  // aim at ~4 CTAs per SM, at least 64 K elements per split.
  const int ctas = cdiv(m, 64) * cdiv(n, 64);
  if (ctas >= 2 * dwc_num_sms() || k < 128) return 1;
  int s = cdiv(4 * dwc_num_sms(), ctas);
  const int smax = cdiv(k, 64);
  if (s > smax) s = smax;
  return s < 1 ? 1 : s;
}

extern "C" int64_t dwc_sgemm_workspace_bytes(int m, int n, int k) {
  const int s = sgemm_splits(m, n, k);
  return s > 1 ? (int64_t)s * m * n * sizeof(float) : 0;
}

static int g_tf32 = 0;
extern "C" void dwc_set_tf32(int on) { g_tf32 = on; }
extern "C" int dwc_get_tf32(void) { return g_tf32; }

extern "C" int dwc_sgemm_ws(int m, int n, int k, float alpha, const void* a, int a_dtype, int64_t a_sm, int64_t a_sk,
                            const float* b, int64_t b_sk, int64_t b_sn, float beta, float* c, int64_t c_sm, int64_t c_sn,
                            const float* bias, int act, float* workspace, int64_t workspace_bytes, dwc_stream_t stream) {
  DWC_CHECK(m > 0 && n > 0 && k > 0, "dwc_sgemm: empty problem (%d,%d,%d)", m, n, k);
  // bf16 product mode: large fp32 GEMMs run on the tensor cores (tcgen05 kind::tf32, dense_tc.cu)
  if (g_tf32 && a_dtype == DWC_F32 && (long long)m * n * k >= (1ll << 22) && act <= 1 &&
      dwc_gemm_tf32_ok(m, n, k, a, a_sm, a_sk, b, b_sk, b_sn, c, c_sm, c_sn))
    return dwc_gemm_tf32(m, n, k, alpha, reinterpret_cast<const float*>(a), a_sm, a_sk, b, b_sk, b_sn, beta, c, c_sm, bias,
                         act, stream);
  cudaStream_t st0 = as_stream(stream);
  if (m >= 192 && n >= 192 && cdiv(m, 128) * cdiv(n, 128) >= dwc_num_sms() / 2) {
    dim3 g128(cdiv(m, 128), cdiv(n, 128));
    if (a_dtype == DWC_F32)
      sgemm128_kernel<float><<<g128, 256, 0, st0>>>(m, n, k, alpha, reinterpret_cast<const float*>(a), a_sm, a_sk, b,
                                                    b_sk, b_sn, beta, c, c_sm, c_sn, bias, act);
    else
      sgemm128_kernel<bf16><<<g128, 256, 0, st0>>>(m, n, k, alpha, reinterpret_cast<const bf16*>(a), a_sm, a_sk, b,
                                                   b_sk, b_sn, beta, c, c_sm, c_sn, bias, act);
    DWC_LAUNCH_CHECK();
    return 0;
  }
  int splits = workspace ? sgemm_splits(m, n, k) : 1;
  if (splits > 1 && workspace_bytes < (int64_t)splits * m * n * (int64_t)sizeof(float)) splits = 1;
  int kper = cdiv(cdiv(k, splits), 16) * 16;
  splits = cdiv(k, kper);
  dim3 grid(cdiv(m, 64), cdiv(n, 64), splits);
  cudaStream_t st = as_stream(stream);
  if (a_dtype == DWC_F32)
    sgemm_kernel<float><<<grid, 256, 0, st>>>(m, n, k, alpha, reinterpret_cast<const float*>(a), a_sm, a_sk, b, b_sk,
                                              b_sn, beta, c, c_sm, c_sn, bias, act, kper, workspace);
  else
    sgemm_kernel<bf16><<<grid, 256, 0, st>>>(m, n, k, alpha, reinterpret_cast<const bf16*>(a), a_sm, a_sk, b, b_sk,
                                             b_sn, beta, c, c_sm, c_sn, bias, act, kper, workspace);
  DWC_LAUNCH_CHECK();
  if (splits > 1) {
    sgemm_splitk_finish_kernel<<<cdiv((long long)m * n, 256), 256, 0, st>>>(m, n, splits, alpha, workspace, beta, c,
                                                                            c_sm, c_sn, bias, act);
    DWC_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int dwc_sgemm(int m, int n, int k, float alpha, const void* a, int a_dtype, int64_t a_sm, int64_t a_sk,
                         const float* b, int64_t b_sk, int64_t b_sn, float beta, float* c, int64_t c_sm, int64_t c_sn,
                         const float* bias, int act, dwc_stream_t stream) {
  return dwc_sgemm_ws(m, n, k, alpha, a, a_dtype, a_sm, a_sk, b, b_sk, b_sn, beta, c, c_sm, c_sn, bias, act, nullptr, 0,
                      stream);
}

__global__ void colsum_kernel(int M, int N, const float* __restrict__ A, long long a_sm, long long a_sn,
                              float* __restrict__ out, int accumulate) {
  // one warp per column
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  float s = 0.f;
  for (int m = lane; m < M; m += 32) s += A[m * a_sm + warp * a_sn];
  s = warp_sum(s);
  if (lane == 0) out[warp] = accumulate ? out[warp] + s : s;
}
extern "C" int dwc_colsum(int m, int n, const float* a, int64_t a_sm, int64_t a_sn, float* out, int accumulate,
                          dwc_stream_t stream) {
  colsum_kernel<<<cdiv((long long)n * 32, 256), 256, 0, as_stream(stream)>>>(m, n, a, a_sm, a_sn, out, accumulate);
  DWC_LAUNCH_CHECK();
  return 0;
}

__global__ void relu_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, float* __restrict__ din,
                                long long count) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    din[i] = out[i] > 0.f ? dout[i] : 0.f;
}
__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                           long long count) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    out[i] = a[i] * b[i];
}
static inline int grid1d(long long count) {
  long long b = (count + 255) / 256;
  return (int)(b > 2368 ? 2368 : (b < 1 ? 1 : b));
}
extern "C" int dwc_relu_bwd(const float* dout, const float* out, float* din, int64_t count, dwc_stream_t stream) {
  relu_bwd_kernel<<<grid1d(count), 256, 0, as_stream(stream)>>>(dout, out, din, count);
  DWC_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwc_mul(const float* a, const float* b, float* out, int64_t count, dwc_stream_t stream) {
  mul_kernel<<<grid1d(count), 256, 0, as_stream(stream)>>>(a, b, out, count);
  DWC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// embedding lookup + dropout mask + style concat:  x[t,b,:] = [emb[tok[b,t]] * mask[t,b,:], style[b,:]]
// ---------------------------------------------------------------------------------------------------
__global__ void embed_concat_fwd_kernel(const int64_t* __restrict__ tok, const float* __restrict__ emb,
                                        const float* __restrict__ style, const float* __restrict__ mask,
                                        float* __restrict__ x, int B, int T, int E, int S) {
  const int D = E + S;
  const long long total = (long long)T * B * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int d = (int)(i % D);
    long long r = i / D;
    int b = (int)(r % B), t = (int)(r / B);
    float v;
    if (d < E) {
      v = emb[tok[(long long)b * T + t] * E + d];
      if (mask) v *= mask[((long long)t * B + b) * E + d];
    } else {
      v = style[(long long)b * S + (d - E)];
    }
    x[i] = v;
  }
}
// dstyle[b] = sum_t dx[t,b,E:]
__global__ void embed_concat_bwd_style_kernel(const float* __restrict__ dx, float* __restrict__ dstyle, int B, int T,
                                              int E, int S) {
  const int D = E + S;
  const long long total = (long long)B * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = E + (int)(i % S);
    const int b = (int)(i / S);
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += dx[((long long)t * B + b) * D + d];
    dstyle[(long long)b * S + (d - E)] = s;
  }
}
// demb[token] += sum over the positions holding that token of dx * mask, WITHOUT atomics: one block per position; the
// block of the FIRST position of a token (in b-major, t-minor order) owns the token's row and adds the contributions of
// all its positions in that fixed order, every other block exits.  Bit-reproducible run to run (float atomics in
// arrival order are not, and Adam turns a one-ulp gradient difference into a full +-lr step for near-zero gradients).
__global__ void __launch_bounds__(128)
    embed_bwd_owner_kernel(const int64_t* __restrict__ tok, const float* __restrict__ dx, const float* __restrict__ mask,
                           float* __restrict__ demb, int B, int T, int E, int S, int pad_idx) {
  extern __shared__ int stok[];                          // all P tokens: the scans below run out of shared memory
  const int P = B * T, p = blockIdx.x;
  for (int q = threadIdx.x; q < P; q += blockDim.x) stok[q] = (int)tok[q];
  __syncthreads();
  const int token = stok[p];
  if (token == pad_idx) return;
  int earlier = 0;
  for (int q = threadIdx.x; q < p; q += blockDim.x) earlier |= (stok[q] == token);
  if (__syncthreads_or(earlier)) return;
  // ordered list of this token's positions (128 candidates per round, compacted with warp ballots)
  int* mlist = stok + P;
  __shared__ int wcount[4], total;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int base = p; base < P; base += 128) {
    const int q = base + threadIdx.x;
    const bool m = q < P && stok[q] == token;
    const unsigned bal = __ballot_sync(0xffffffffu, m);
    if (lane == 0) wcount[warp] = __popc(bal);
    __syncthreads();
    int off = total;
    for (int w2 = 0; w2 < warp; ++w2) off += wcount[w2];
    if (m) mlist[off + __popc(bal & ((1u << lane) - 1u))] = q;
    __syncthreads();
    if (threadIdx.x == 0) total += wcount[0] + wcount[1] + wcount[2] + wcount[3];
    __syncthreads();
  }
  const int cnt = total;
  const int D = E + S;
  for (int d0 = 0; d0 < E; d0 += 4 * 128) {            // 4 dims per thread, positions in fixed (b, t) order
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < cnt; ++i) {
      const int q = mlist[i];
      const int b = q / T, t = q - b * T;
      const long long row = (long long)t * B + b;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = d0 + j * 128 + threadIdx.x;
        if (d < E) {
          float g = dx[row * D + d];
          if (mask) g *= mask[row * E + d];
          acc[j] += g;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = d0 + j * 128 + threadIdx.x;
      if (d < E) demb[(long long)token * E + d] += acc[j];
    }
  }
}
extern "C" int dwc_embed_concat_fwd(const int64_t* tokens, const float* emb, const float* style, const float* mask,
                                    float* x, int b, int t, int e, int s, dwc_stream_t stream) {
  embed_concat_fwd_kernel<<<grid1d((long long)t * b * (e + s)), 256, 0, as_stream(stream)>>>(tokens, emb, style, mask,
                                                                                               x, b, t, e, s);
  DWC_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwc_embed_concat_bwd(const int64_t* tokens, const float* dx, const float* mask, float* demb,
                                    float* dstyle, int b, int t, int e, int s, int pad_idx, dwc_stream_t stream) {
  embed_concat_bwd_style_kernel<<<grid1d((long long)b * s), 256, 0, as_stream(stream)>>>(dx, dstyle, b, t, e, s);
  DWC_LAUNCH_CHECK();
  if (demb) {
    DWC_CHECK((long long)b * t * 8 <= 48 * 1024, "dwc_embed_concat_bwd: %d x %d tokens exceed the shared-memory lists", b, t);
    embed_bwd_owner_kernel<<<b * t, 128, (size_t)b * t * 8, as_stream(stream)>>>(tokens, dx, mask, demb, b, t, e, s,
                                                                                 pad_idx);
    DWC_LAUNCH_CHECK();
  }
  return 0;
}

// batched 2-D transpose: dst[b][c][r] = src[b][r][c]
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int Cc) {
  __shared__ float tile[32][33];
  const float* s = src + (long long)blockIdx.z * R * Cc;
  float* d = dst + (long long)blockIdx.z * R * Cc;
  int c = blockIdx.x * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int r = blockIdx.y * 32 + i;
    if (r < R && c < Cc) tile[i][threadIdx.x] = s[(long long)r * Cc + c];
  }
  __syncthreads();
  int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int c2 = blockIdx.x * 32 + i;
    if (r2 < R && c2 < Cc) d[(long long)c2 * R + r2] = tile[threadIdx.x][i];
  }
}
extern "C" int dwc_transpose(const float* src, float* dst, int batch, int rows, int cols, dwc_stream_t stream) {
  dim3 grid(cdiv(cols, 32), cdiv(rows, 32), batch), block(32, 8);
  transpose_kernel<<<grid, block, 0, as_stream(stream)>>>(src, dst, rows, cols);
  DWC_LAUNCH_CHECK();
  return 0;
}
