// Text-encoder recurrence (networks/networks_v2.py:225-233: nn.LSTM over a packed sequence) as ONE persistent
// kernel per layer and pass.  Both directions and all time steps run inside a single launch:
//
//   grid (NC, 2): direction d = blockIdx.y, CTA x owns UPC hidden units j0..j0+UPC-1 of that direction and keeps
//   its slice of the recurrent weights in shared memory for the whole sequence.  Per time step every CTA
//     forward : reads h_{t-1} of its direction (B x H floats, written by the NC sibling CTAs) from L2,
//               computes its 4*UPC gate rows for all samples, applies the cell update for its units;
//     backward: reads the gate gradients of the step processed just before (B x 4H floats) from L2,
//               computes W_hh^T dgates for its units, applies the point-wise backward for its units;
//   then the NC CTAs of the direction meet at a device-side barrier (monotonic counter in global memory;
//   the kernel is launched cooperatively so that all CTAs are co-resident).
//
// The loop runs over T_eff = max(lens) steps only; per-sample lengths give pack_padded_sequence semantics
// (state frozen, output zero at t >= len[b]).  Rows t >= T_eff of out / dgates are zero-filled.
#include "common.cuh"

namespace {

constexpr int UPC = 5;          // hidden units per CTA
constexpr int ROWS = 4 * UPC;   // gate rows per CTA
constexpr int NTHR = 256;
constexpr int NWARP = NTHR / 32;
constexpr int BCH = 16;         // samples per compute chunk

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all CTAs of one direction: arrive + wait until `target` arrivals have been counted
__device__ __forceinline__ void dir_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned spins = 0;
    while (ld_acquire(counter) < target) {
      if (++spins > (1u << 24)) asm volatile("trap;");   // a lost arrival becomes a launch failure, not a hang
    }
    __threadfence();
  }
  __syncthreads();
}

struct FwdP {
  int T, B, H, NC;
  const float* xproj;       // [T,B,2,4H]
  const float* whh;         // [2,4H,H]
  const long long* lens;    // [B]
  float* hbuf;              // [2 parity][2 dir][B][H] exchange buffer
  float* out;               // [T,B,2H] or null
  float* gates_save;        // [T,B,2,4H] or null
  float* c_save;            // [T,B,2,H] or null
  float* h_final;           // [2,B,H]
  float* c_final;           // [2,B,H]
  unsigned* bar;            // [2], zeroed before launch
};

__global__ void __launch_bounds__(NTHR) lstm_layer_fwd_kernel(const FwdP p) {
  extern __shared__ float sm[];
  const int H = p.H, B = p.B, T = p.T;
  const int HP = H + 1;
  float* Wt = sm;                               // [H][ROWS]   (k-major: a thread reads consecutive rows)
  float* hs = Wt + H * ROWS;                    // [B][HP]
  float* red = hs + B * HP;                     // [NWARP][ROWS][BCH]
  float* cst = red + NWARP * ROWS * BCH;        // [B][UPC] cell state of the own units
  float* hst = cst + B * UPC;                   // [B][UPC] hidden state of the own units
  __shared__ int s_teff;

  const int dir = blockIdx.y;
  const int j0 = blockIdx.x * UPC;
  const int nu = min(UPC, H - j0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    long long m = 0;
    for (int b = 0; b < B; ++b) m = max(m, p.lens[b]);
    s_teff = (int)min((long long)T, m);
  }
  // recurrent weight slice: row r = q*UPC + u  <->  whh[dir][q*H + j0 + u][:]
  for (int i = tid; i < ROWS * H; i += NTHR) {
    const int r = i / H, k = i - r * H;
    const int q = r / UPC, u = r - q * UPC;
    Wt[k * ROWS + r] = (u < nu) ? p.whh[((long long)dir * 4 * H + q * H + j0 + u) * H + k] : 0.f;
  }
  for (int i = tid; i < B * UPC; i += NTHR) { cst[i] = 0.f; hst[i] = 0.f; }
  __syncthreads();
  const int Teff = s_teff;
  // zero the rows of `out` past the longest sequence (they feed the next layer's GEMMs)
  if (p.out) {
    for (int i = tid; i < (T - Teff) * B * nu; i += NTHR) {
      const int u = i % nu, r = i / nu;
      const int b = r % B, t = Teff + r / B;
      p.out[((long long)t * B + b) * 2 * H + dir * H + j0 + u] = 0.f;
    }
  }

  const int KW = (H + NWARP - 1) / NWARP;
  const int bl = lane & 15, rg = lane >> 4;     // compute role: sample within the chunk, row group (ROWS/2 rows)
  constexpr int RPG = ROWS / 2;

  for (int s = 0; s < Teff; ++s) {
    const int t = dir == 0 ? s : Teff - 1 - s;
    for (int b0 = 0; b0 < B; b0 += BCH) {
      const int nb = min(BCH, B - b0);
      // pre-activations from the input projection: fetched early, consumed after the mat-vec
      float xp[4] = {0.f, 0.f, 0.f, 0.f};
      const int eu = tid % UPC, eb = tid / UPC;     // point-wise role: (unit, sample) for tid < UPC*nb
      const bool ev = tid < UPC * nb && eu < nu;
      if (ev) {
        const float* xr = p.xproj + (((long long)t * B + b0 + eb) * 2 + dir) * 4 * H + j0 + eu;
#pragma unroll
        for (int q = 0; q < 4; ++q) xp[q] = __ldg(xr + q * H);
      }
      float acc[RPG];
#pragma unroll
      for (int i = 0; i < RPG; ++i) acc[i] = 0.f;
      if (s > 0) {
        if (b0 == 0) {
          // h_{t-1} of this direction, all samples (written by the sibling CTAs before the last barrier)
          const float* hsrc = p.hbuf + ((long long)((s - 1) & 1) * 2 + dir) * B * H;
          if ((H & 3) == 0) {
            const int H4 = H >> 2;
            for (int i = tid; i < B * H4; i += NTHR) {
              const int b = i / H4, k4 = i - b * H4;
              const float4 v = __ldcg(reinterpret_cast<const float4*>(hsrc + (long long)b * H) + k4);
              float* d = hs + b * HP + 4 * k4;
              d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            }
          } else {
            for (int i = tid; i < B * H; i += NTHR) {
              const int b = i / H, k = i - b * H;
              hs[b * HP + k] = __ldcg(hsrc + i);
            }
          }
          __syncthreads();
        }
        const int k0 = warp * KW, k1 = min(H, k0 + KW);
        if (bl < nb) {
          const float* hrow = hs + (b0 + bl) * HP;
          for (int k = k0; k < k1; ++k) {
            const float hv = hrow[k];
            const float2* wr = reinterpret_cast<const float2*>(Wt + k * ROWS + rg * RPG);
#pragma unroll
            for (int i = 0; i < RPG / 2; ++i) {
              const float2 w2 = wr[i];
              acc[2 * i] = fmaf(hv, w2.x, acc[2 * i]);
              acc[2 * i + 1] = fmaf(hv, w2.y, acc[2 * i + 1]);
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < RPG; ++i) red[(warp * ROWS + rg * RPG + i) * BCH + bl] = acc[i];
      __syncthreads();
      if (ev) {
        float pre[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v = xp[q];
#pragma unroll
          for (int w = 0; w < NWARP; ++w) v += red[(w * ROWS + q * UPC + eu) * BCH + eb];
          pre[q] = v;
        }
        const int bb = b0 + eb, jj = j0 + eu;
        const float gi = sigm(pre[0]), gf = sigm(pre[1]), gg = tanhf(pre[2]), go = sigm(pre[3]);
        const float cprev = cst[bb * UPC + eu], hprev = hst[bb * UPC + eu];
        const bool active = (long long)t < p.lens[bb];
        const float cn = active ? gf * cprev + gi * gg : cprev;
        const float hn = active ? go * tanhf(cn) : hprev;
        cst[bb * UPC + eu] = cn;
        hst[bb * UPC + eu] = hn;
        p.hbuf[(((long long)(s & 1) * 2 + dir) * B + bb) * H + jj] = hn;
        if (p.out) p.out[((long long)t * B + bb) * 2 * H + dir * H + jj] = active ? hn : 0.f;
        if (p.gates_save) {
          float* g4 = p.gates_save + (((long long)t * B + bb) * 2 + dir) * 4 * H;
          g4[jj] = gi; g4[H + jj] = gf; g4[2 * H + jj] = gg; g4[3 * H + jj] = go;
          p.c_save[(((long long)t * B + bb) * 2 + dir) * H + jj] = cn;
        }
      }
      __syncthreads();
    }
    if (s + 1 < Teff) dir_barrier(p.bar + dir, (unsigned)(s + 1) * p.NC);
  }
  for (int i = tid; i < B * nu; i += NTHR) {
    const int u = i % nu, b = i / nu;
    p.h_final[((long long)dir * B + b) * H + j0 + u] = hst[b * UPC + u];
    p.c_final[((long long)dir * B + b) * H + j0 + u] = cst[b * UPC + u];
  }
}

struct BwdP {
  int T, B, H, NC;
  const float* whh;         // [2,4H,H]
  const long long* lens;
  const float* dout;        // [T,B,2H] or null
  const float* gates_save;  // [T,B,2,4H]
  const float* c_save;      // [T,B,2,H]
  const float* dh0;         // [2,B,H] gradient of the final hidden state
  const float* dc0;         // [2,B,H] gradient of the final cell state
  float* dgates;            // [T,B,2,4H]
  unsigned* bar;
};

__global__ void __launch_bounds__(NTHR) lstm_layer_bwd_kernel(const BwdP p) {
  extern __shared__ float sm[];
  const int H = p.H, B = p.B, T = p.T;
  const int G = 4 * H, GP = G + 1;
  float* Wc = sm;                               // [4H][8]   Wc[r][u] = whh[dir][r][j0+u]
  float* dgs = Wc + G * 8;                      // [BCH][GP] gate gradients of the later step (one sample chunk)
  float* red = dgs + BCH * GP;                  // [2*NWARP][UPC][BCH]
  float* dhs = red + 2 * NWARP * UPC * BCH;     // [B][UPC]
  float* dcs = dhs + B * UPC;                   // [B][UPC]
  __shared__ int s_teff;

  const int dir = blockIdx.y;
  const int j0 = blockIdx.x * UPC;
  const int nu = min(UPC, H - j0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    long long m = 0;
    for (int b = 0; b < B; ++b) m = max(m, p.lens[b]);
    s_teff = (int)min((long long)T, m);
  }
  for (int i = tid; i < G * 8; i += NTHR) {
    const int r = i >> 3, u = i & 7;
    Wc[i] = (u < nu) ? p.whh[((long long)dir * G + r) * H + j0 + u] : 0.f;
  }
  for (int i = tid; i < B * UPC; i += NTHR) {
    const int u = i % UPC, b = i / UPC;
    const bool v = u < nu;
    dhs[i] = v ? p.dh0[((long long)dir * B + b) * H + j0 + u] : 0.f;
    dcs[i] = v ? p.dc0[((long long)dir * B + b) * H + j0 + u] : 0.f;
  }
  __syncthreads();
  const int Teff = s_teff;
  for (int i = tid; i < (T - Teff) * B * 4 * nu; i += NTHR) {
    const int u = i % nu;
    int r = i / nu;
    const int q = r & 3;
    r >>= 2;
    const int b = r % B, t = Teff + r / B;
    p.dgates[(((long long)t * B + b) * 2 + dir) * G + q * H + j0 + u] = 0.f;
  }

  const int bl = lane & 15, kh = lane >> 4;     // compute role: sample within the chunk, half of the warp's k range
  const int KW = (G + 2 * NWARP - 1) / (2 * NWARP);

  for (int s = 0; s < Teff; ++s) {
    const int t = dir == 0 ? Teff - 1 - s : s;
    const int t_later = dir == 0 ? t + 1 : t - 1;
    const int t_prev = dir == 0 ? t - 1 : t + 1;
    for (int b0 = 0; b0 < B; b0 += BCH) {
      const int nb = min(BCH, B - b0);
      float acc[UPC];
#pragma unroll
      for (int u = 0; u < UPC; ++u) acc[u] = 0.f;
      if (s > 0) {
        const float* src = p.dgates + (((long long)t_later * B + b0) * 2 + dir) * G;   // sample stride 2*G
        {
          const int G4 = G >> 2;                       // G = 4H is always a multiple of 4; rows are 16-byte aligned
          for (int i = tid; i < nb * G4; i += NTHR) {
            const int b = i / G4, g4 = i - b * G4;
            const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (long long)b * 2 * G) + g4);
            float* d = dgs + b * GP + 4 * g4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
          }
        }
        __syncthreads();
        const int k0 = (warp * 2 + kh) * KW, k1 = min(G, k0 + KW);
        if (bl < nb) {
          const float* drow = dgs + bl * GP;
          for (int k = k0; k < k1; ++k) {
            const float dv = drow[k];
            const float4 w4 = *reinterpret_cast<const float4*>(Wc + k * 8);
            const float w5 = Wc[k * 8 + 4];
            acc[0] = fmaf(dv, w4.x, acc[0]);
            acc[1] = fmaf(dv, w4.y, acc[1]);
            acc[2] = fmaf(dv, w4.z, acc[2]);
            acc[3] = fmaf(dv, w4.w, acc[3]);
            acc[4] = fmaf(dv, w5, acc[4]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UPC; ++u) red[((warp * 2 + kh) * UPC + u) * BCH + bl] = acc[u];
      __syncthreads();
      const int eu = tid % UPC, eb = tid / UPC;
      if (tid < UPC * nb && eu < nu) {
        float rec = 0.f;
#pragma unroll
        for (int w = 0; w < 2 * NWARP; ++w) rec += red[(w * UPC + eu) * BCH + eb];
        const int bb = b0 + eb, jj = j0 + eu;
        const bool active = (long long)t < p.lens[bb];
        float dh = dhs[bb * UPC + eu] + rec;
        const float dc = dcs[bb * UPC + eu];
        float* dg = p.dgates + (((long long)t * B + bb) * 2 + dir) * G;
        if (active) {
          if (p.dout) dh += p.dout[((long long)t * B + bb) * 2 * H + dir * H + jj];
          const float* g4 = p.gates_save + (((long long)t * B + bb) * 2 + dir) * G;
          const float gi = g4[jj], gf = g4[H + jj], gg = g4[2 * H + jj], go = g4[3 * H + jj];
          const float cn = p.c_save[(((long long)t * B + bb) * 2 + dir) * H + jj];
          float cprev = 0.f;
          if (t_prev >= 0 && t_prev < Teff) cprev = p.c_save[(((long long)t_prev * B + bb) * 2 + dir) * H + jj];
          const float tc = tanhf(cn);
          const float dco = dc + dh * go * (1.f - tc * tc);
          dg[jj] = dco * gg * gi * (1.f - gi);
          dg[H + jj] = dco * cprev * gf * (1.f - gf);
          dg[2 * H + jj] = dco * gi * (1.f - gg * gg);
          dg[3 * H + jj] = dh * tc * go * (1.f - go);
          dcs[bb * UPC + eu] = dco * gf;
          dhs[bb * UPC + eu] = 0.f;          // the recurrent part arrives through dgates at the next step
        } else {
          dg[jj] = 0.f; dg[H + jj] = 0.f; dg[2 * H + jj] = 0.f; dg[3 * H + jj] = 0.f;
          dhs[bb * UPC + eu] = dh;           // state passes through a padded step unchanged
        }
      }
      __syncthreads();
    }
    if (s + 1 < Teff) dir_barrier(p.bar + dir, (unsigned)(s + 1) * p.NC);
  }
}

int check_coop(const void* fn, int nthreads, size_t smem, int blocks) {
  int dev = 0, sms = 0, per_sm = 0, coop = 0;
  DWC_CUDA(cudaGetDevice(&dev));
  DWC_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  DWC_CHECK(coop, "dwc_lstm: device does not support cooperative launches");
  DWC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DWC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, nthreads, smem));
  DWC_CHECK(per_sm * sms >= blocks, "dwc_lstm: %d CTAs cannot be co-resident (%d SMs x %d)", blocks, sms, per_sm);
  return 0;
}

}  // namespace

extern "C" int64_t dwc_lstm_workspace_bytes(int b, int h) {
  return 256 + (int64_t)2 * 2 * b * h * sizeof(float);
}

extern "C" int dwc_lstm_layer_fwd(int t_total, int b, int h, const float* xproj, const float* whh, const int64_t* lens,
                                  float* out, float* gates_save, float* c_save, float* h_final, float* c_final,
                                  void* workspace, dwc_stream_t stream) {
  DWC_CHECK(t_total > 0 && b > 0 && h > 0, "dwc_lstm_layer_fwd: empty problem");
  DWC_CHECK((gates_save == nullptr) == (c_save == nullptr), "dwc_lstm_layer_fwd: gates_save and c_save go together");
  FwdP p;
  p.T = t_total; p.B = b; p.H = h; p.NC = cdiv(h, UPC);
  p.xproj = xproj; p.whh = whh; p.lens = reinterpret_cast<const long long*>(lens);
  p.bar = reinterpret_cast<unsigned*>(workspace);
  p.hbuf = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
  p.out = out; p.gates_save = gates_save; p.c_save = c_save; p.h_final = h_final; p.c_final = c_final;
  const size_t smem = ((size_t)h * ROWS + (size_t)b * (h + 1) + NWARP * ROWS * BCH + 2 * (size_t)b * UPC) * sizeof(float);
  DWC_CHECK(smem <= 220 * 1024, "dwc_lstm_layer_fwd: batch %d x hidden %d does not fit shared memory", b, h);
  DWC_CUDA(cudaFuncSetAttribute(lstm_layer_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (check_coop((const void*)lstm_layer_fwd_kernel, NTHR, smem, p.NC * 2)) return 1;
  cudaStream_t st = as_stream(stream);
  DWC_CUDA(cudaMemsetAsync(workspace, 0, 256, st));
  void* args[] = {&p};
  DWC_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_layer_fwd_kernel, dim3(p.NC, 2), dim3(NTHR), args, smem, st));
  return 0;
}

extern "C" int dwc_lstm_layer_bwd(int t_total, int b, int h, const float* whh, const int64_t* lens, const float* dout,
                                  const float* gates_save, const float* c_save, const float* dh_final,
                                  const float* dc_final, float* dgates, void* workspace, dwc_stream_t stream) {
  DWC_CHECK(t_total > 0 && b > 0 && h > 0, "dwc_lstm_layer_bwd: empty problem");
  BwdP p;
  p.T = t_total; p.B = b; p.H = h; p.NC = cdiv(h, UPC);
  p.whh = whh; p.lens = reinterpret_cast<const long long*>(lens); p.dout = dout;
  p.gates_save = gates_save; p.c_save = c_save; p.dh0 = dh_final; p.dc0 = dc_final; p.dgates = dgates;
  p.bar = reinterpret_cast<unsigned*>(workspace);
  const size_t smem = ((size_t)4 * h * 8 + (size_t)BCH * (4 * h + 1) + 2 * NWARP * UPC * BCH + 2 * (size_t)b * UPC) *
                      sizeof(float);
  DWC_CHECK(smem <= 220 * 1024, "dwc_lstm_layer_bwd: hidden %d does not fit shared memory", h);
  DWC_CUDA(cudaFuncSetAttribute(lstm_layer_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (check_coop((const void*)lstm_layer_bwd_kernel, NTHR, smem, p.NC * 2)) return 1;
  cudaStream_t st = as_stream(stream);
  DWC_CUDA(cudaMemsetAsync(workspace, 0, 256, st));
  void* args[] = {&p};
  DWC_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_layer_bwd_kernel, dim3(p.NC, 2), dim3(NTHR), args, smem, st));
  return 0;
}
