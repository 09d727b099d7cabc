// Small dense algebra and the text encoder: strided fp32 GEMM (style MLP, mapping network, heads,
// LSTM input projections and weight gradients), embedding + style concat, fused LSTM time steps.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------
// C[m,n] = act(alpha * sum_k A[m,k] B[k,n] + bias[n] + beta*C[m,n]); 64x64 tile, 256 threads x (4x4)
// ---------------------------------------------------------------------------------------------------
template <typename TA>
__global__ void __launch_bounds__(256)
    sgemm_kernel(int M, int N, int K, float alpha, const TA* __restrict__ A, long long a_sm, long long a_sk,
                 const float* __restrict__ B, long long b_sk, long long b_sn, float beta, float* __restrict__ C,
                 long long c_sm, long long c_sn, const float* __restrict__ bias, int act) {
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
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int e = tid + j * 256;   // 0..1023
      int mm, kk;
      if (a_kfast) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? to_f<TA>(A[gm * a_sm + gk * a_sk]) : 0.f;
      int nn, k2;
      if (b_nfast) { nn = e & 63; k2 = e >> 6; } else { k2 = e & 15; nn = e >> 4; }
      int gn = n0 + nn, gk2 = k0 + k2;
      Bs[k2][nn] = (gn < N && gk2 < K) ? B[gk2 * b_sk + gn * b_sn] : 0.f;
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
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float* o = C + m * c_sm + n * c_sn;
      float v = alpha * acc[i][j] + (bias ? bias[n] : 0.f) + (beta != 0.f ? beta * *o : 0.f);
      if (act == 1) v = v > 0.f ? v : 0.f;
      *o = v;
    }
  }
}

extern "C" int dwc_sgemm(int m, int n, int k, float alpha, const void* a, int a_dtype, int64_t a_sm, int64_t a_sk,
                         const float* b, int64_t b_sk, int64_t b_sn, float beta, float* c, int64_t c_sm, int64_t c_sn,
                         const float* bias, int act, dwc_stream_t stream) {
  DWC_CHECK(m > 0 && n > 0 && k > 0, "dwc_sgemm: empty problem (%d,%d,%d)", m, n, k);
  dim3 grid(cdiv(m, 64), cdiv(n, 64));
  if (a_dtype == DWC_F32)
    sgemm_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(m, n, k, alpha, reinterpret_cast<const float*>(a), a_sm,
                                                              a_sk, b, b_sk, b_sn, beta, c, c_sm, c_sn, bias, act);
  else
    sgemm_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>(m, n, k, alpha, reinterpret_cast<const bf16*>(a), a_sm,
                                                             a_sk, b, b_sk, b_sn, beta, c, c_sm, c_sn, bias, act);
  DWC_LAUNCH_CHECK();
  return 0;
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
// demb[token] += dx (atomics; token rows collide), dstyle[b] = sum_t dx[t,b,E:]
__global__ void embed_concat_bwd_kernel(const int64_t* __restrict__ tok, const float* __restrict__ dx,
                                        const float* __restrict__ mask, float* __restrict__ demb,
                                        float* __restrict__ dstyle, int B, int T, int E, int S, int pad_idx) {
  const int D = E + S;
  const long long total = (long long)B * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int d = (int)(i % D);
    int b = (int)(i / D);
    if (d >= E) {
      float s = 0.f;
      for (int t = 0; t < T; ++t) s += dx[((long long)t * B + b) * D + d];
      dstyle[(long long)b * S + (d - E)] = s;
    } else if (demb) {
      for (int t = 0; t < T; ++t) {
        long long token = tok[(long long)b * T + t];
        if (token == pad_idx) continue;
        float g = dx[((long long)t * B + b) * D + d];
        if (mask) g *= mask[((long long)t * B + b) * E + d];
        atomicAdd(demb + token * E + d, g);
      }
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
  embed_concat_bwd_kernel<<<grid1d((long long)b * (e + s)), 256, 0, as_stream(stream)>>>(tokens, dx, mask, demb, dstyle,
                                                                                          b, t, e, s, pad_idx);
  DWC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// LSTM time step, both directions.  grid (ceil(H/32), 2 dirs); block 128 = 4 gates x 32 hidden units.
// Thread (gate q, unit j) accumulates its gate pre-activation for up to 16 samples at a time:
//   forward : sum_k h_prev[b,k] * WhhT[k, q*H+j]      (WhhT = transposed recurrent weights: coalesced over j)
//   backward: sum_i dgates_later[b, q*H+i] * Whh[q*H+i, j]   (row-major Whh: coalesced over j), summed over q
// then the 128 threads share the point-wise cell update of the 32 units x B samples.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

constexpr int LSTM_BC = 16;     // samples per register chunk

__global__ void __launch_bounds__(128)
    lstm_step_fwd_kernel(int step, int T, int B, int H, const float* __restrict__ xproj, const float* __restrict__ whh_t,
                         const int64_t* __restrict__ lens, const float* __restrict__ h_in, const float* __restrict__ c_in,
                         float* __restrict__ h_out, float* __restrict__ c_out, float* __restrict__ out,
                         float* __restrict__ gates_save, float* __restrict__ c_save) {
  extern __shared__ float sm[];
  float* hs = sm;                                  // [LSTM_BC][H]
  float* gs = sm + LSTM_BC * H;                    // [4][32][LSTM_BC + 1]
  const int dir = blockIdx.y;
  const int t = dir == 0 ? step : T - 1 - step;
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 32 + lane;
  const bool jv = j < H;
  const float* wt = whh_t + (long long)dir * H * 4 * H + (long long)q * H + j;   // + k*4H
  for (int b0 = 0; b0 < B; b0 += LSTM_BC) {
    const int nb = min(LSTM_BC, B - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * H; i += blockDim.x) hs[i] = h_in[((long long)dir * B + b0) * H + i];
    __syncthreads();
    float acc[LSTM_BC];
#pragma unroll
    for (int b = 0; b < LSTM_BC; ++b) acc[b] = 0.f;
    if (jv) {
      for (int k = 0; k < H; ++k) {
        const float w = wt[(long long)k * 4 * H];
#pragma unroll
        for (int b = 0; b < LSTM_BC; ++b) acc[b] = fmaf(hs[b * H + k], w, acc[b]);   // rows >= nb hold stale data, unused
      }
    }
#pragma unroll
    for (int b = 0; b < LSTM_BC; ++b) {
      float pre = 0.f;
      if (jv && b < nb) pre = acc[b] + xproj[((((long long)t * B + b0 + b) * 2 + dir) * 4 + q) * H + j];
      gs[(q * 32 + lane) * (LSTM_BC + 1) + b] = pre;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * nb; e += blockDim.x) {
      const int l2 = e & 31, b = e >> 5;
      const int jj = blockIdx.x * 32 + l2;
      if (jj >= H) continue;
      const int bb = b0 + b;
      const float gi = sigm(gs[(0 * 32 + l2) * (LSTM_BC + 1) + b]), gf = sigm(gs[(1 * 32 + l2) * (LSTM_BC + 1) + b]);
      const float gg = tanhf(gs[(2 * 32 + l2) * (LSTM_BC + 1) + b]), go = sigm(gs[(3 * 32 + l2) * (LSTM_BC + 1) + b]);
      const long long st = ((long long)dir * B + bb) * H + jj;
      const float hprev = h_in[st], cprev = c_in[st];
      const bool active = (long long)t < lens[bb];
      const float cn = gf * cprev + gi * gg;
      const float hn = go * tanhf(cn);
      h_out[st] = active ? hn : hprev;
      c_out[st] = active ? cn : cprev;
      if (out) out[((long long)t * B + bb) * 2 * H + dir * H + jj] = active ? hn : 0.f;
      if (gates_save) {
        float* g4 = gates_save + (((long long)t * B + bb) * 2 + dir) * 4 * H;
        g4[jj] = gi; g4[H + jj] = gf; g4[2 * H + jj] = gg; g4[3 * H + jj] = go;
        c_save[(((long long)t * B + bb) * 2 + dir) * H + jj] = active ? cn : cprev;
      }
    }
  }
}

extern "C" int dwc_lstm_step_fwd(int step, int t_total, int b, int h, const float* xproj, const float* whh_t,
                                 const int64_t* lens, const float* h_in, const float* c_in, float* h_out, float* c_out,
                                 float* out, float* gates_save, float* c_save, dwc_stream_t stream) {
  size_t smem = ((size_t)LSTM_BC * h + 4 * 32 * (LSTM_BC + 1)) * sizeof(float);
  DWC_CHECK(smem <= 48 * 1024, "dwc_lstm_step_fwd: hidden size too large");
  dim3 grid(cdiv(h, 32), 2);
  lstm_step_fwd_kernel<<<grid, 128, smem, as_stream(stream)>>>(step, t_total, b, h, xproj, whh_t, lens, h_in, c_in,
                                                               h_out, c_out, out, gates_save, c_save);
  DWC_LAUNCH_CHECK();
  return 0;
}

// Backward of one step.  For direction d at time t (processed in reverse order of the forward):
//   dh_t = dh_state (+ W_hh^T dgates of the step processed just before, folded in here) + dout[t]
__global__ void __launch_bounds__(128)
    lstm_step_bwd_kernel(int step, int T, int B, int H, const float* __restrict__ whh, const int64_t* __restrict__ lens,
                         const float* __restrict__ dout, const float* __restrict__ gates_save,
                         const float* __restrict__ c_save, const float* __restrict__ dh_in, const float* __restrict__ dc_in,
                         float* __restrict__ dh_out, float* __restrict__ dc_out, float* __restrict__ dgates) {
  extern __shared__ float sm[];
  float* dgs = sm;                                 // [LSTM_BC][4H] gate gradients of the later step
  float* part = sm + LSTM_BC * 4 * H;              // [4][32][LSTM_BC + 1]
  const int dir = blockIdx.y;
  const int t = dir == 0 ? T - 1 - step : step;
  const int t_later = dir == 0 ? t + 1 : t - 1;
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 32 + lane;
  const bool jv = j < H;
  const bool has_later = step > 0;
  const float* w = whh + (long long)dir * 4 * H * H + (long long)q * H * H + j;   // + i*H
  for (int b0 = 0; b0 < B; b0 += LSTM_BC) {
    const int nb = min(LSTM_BC, B - b0);
    float acc[LSTM_BC];
#pragma unroll
    for (int b = 0; b < LSTM_BC; ++b) acc[b] = 0.f;
    __syncthreads();
    if (has_later) {
      for (int i = threadIdx.x; i < nb * 4 * H; i += blockDim.x) {
        int b = i / (4 * H), g = i - b * 4 * H;
        dgs[b * 4 * H + g] = dgates[(((long long)t_later * B + b0 + b) * 2 + dir) * 4 * H + g];
      }
      __syncthreads();
      if (jv) {
        for (int i = 0; i < H; ++i) {
          const float wv = w[(long long)i * H];
#pragma unroll
          for (int b = 0; b < LSTM_BC; ++b) acc[b] = fmaf(dgs[b * 4 * H + q * H + i], wv, acc[b]);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < LSTM_BC; ++b) part[(q * 32 + lane) * (LSTM_BC + 1) + b] = acc[b];
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * nb; e += blockDim.x) {
      const int l2 = e & 31, b = e >> 5;
      const int jj = blockIdx.x * 32 + l2;
      if (jj >= H) continue;
      const int bb = b0 + b;
      const float rec = part[(0 * 32 + l2) * (LSTM_BC + 1) + b] + part[(1 * 32 + l2) * (LSTM_BC + 1) + b] +
                        part[(2 * 32 + l2) * (LSTM_BC + 1) + b] + part[(3 * 32 + l2) * (LSTM_BC + 1) + b];
      const long long st = ((long long)dir * B + bb) * H + jj;
      const bool active = (long long)t < lens[bb];
      float dh = dh_in[st] + rec;
      const float dc = dc_in[st];
      float* dg = dgates + (((long long)t * B + bb) * 2 + dir) * 4 * H;
      if (active) {
        if (dout) dh += dout[((long long)t * B + bb) * 2 * H + dir * H + jj];
        const float* g4 = gates_save + (((long long)t * B + bb) * 2 + dir) * 4 * H;
        const float gi = g4[jj], gf = g4[H + jj], gg = g4[2 * H + jj], go = g4[3 * H + jj];
        const float cn = c_save[(((long long)t * B + bb) * 2 + dir) * H + jj];
        const int t_prev = dir == 0 ? t - 1 : t + 1;
        float cprev = 0.f;
        if (t_prev >= 0 && t_prev < T) cprev = c_save[(((long long)t_prev * B + bb) * 2 + dir) * H + jj];
        const float tc = tanhf(cn);
        const float dco = dc + dh * go * (1.f - tc * tc);
        dg[jj] = dco * gg * gi * (1.f - gi);
        dg[H + jj] = dco * cprev * gf * (1.f - gf);
        dg[2 * H + jj] = dco * gi * (1.f - gg * gg);
        dg[3 * H + jj] = dh * tc * go * (1.f - go);
        dc_out[st] = dco * gf;
        dh_out[st] = 0.f;           // the recurrent part is added by the next launch from dgates
      } else {
        dg[jj] = 0.f; dg[H + jj] = 0.f; dg[2 * H + jj] = 0.f; dg[3 * H + jj] = 0.f;
        dc_out[st] = dc;
        dh_out[st] = dh;            // state passes through a padded step unchanged
      }
    }
  }
}

extern "C" int dwc_lstm_step_bwd(int step, int t_total, int b, int h, const float* whh, const int64_t* lens,
                                 const float* dout, const float* gates_save, const float* c_save, const float* dh_in,
                                 const float* dc_in, float* dh_out, float* dc_out, float* dgates, dwc_stream_t stream) {
  size_t smem = ((size_t)LSTM_BC * 4 * h + 4 * 32 * (LSTM_BC + 1)) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    DWC_CUDA(cudaFuncSetAttribute(lstm_step_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  DWC_CHECK(smem <= 200 * 1024, "dwc_lstm_step_bwd: hidden size too large for shared memory");
  dim3 grid(cdiv(h, 32), 2);
  lstm_step_bwd_kernel<<<grid, 128, smem, as_stream(stream)>>>(step, t_total, b, h, whh, lens, dout, gates_save, c_save,
                                                               dh_in, dc_in, dh_out, dc_out, dgates);
  DWC_LAUNCH_CHECK();
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
