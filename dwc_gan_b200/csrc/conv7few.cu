// conv7few: 7x7 stride-1 "valid" convolution from 64 channels to <= 4 channels on tcgen05 - the decoder heads
// (64 -> 3+1, networks_v2.py:162-169) and the data gradient of a first encoder convolution (64 -> 3 image channels,
// the transpose of networks_v2.py:52 / networks.py:442).
//
// The generic tap-by-tap kernel pads the 4 output channels to N = 16 and issues 49 taps x 4 K-steps of M = 128 MMAs
// per 128 pixels; an M = 128 MMA costs >= 64 cycles of A-operand shared-memory read whatever N is, so those layers
// ran at 4-6 % of their bound.  Here the GEMM is re-associated:
//   D[pixel (y, x)][kx*4 + o] = sum_{ky, i}  X[y + ky][x][i] * W[o][ky][kx][i]        (M = 128, N = 32, K = 7 * 64)
//   out[y][x][o]              = sum_{kx}     D[(y, x + kx)][kx*4 + o]
// i.e. the vertical taps are folded into K (the A operand of tap ky is the SAME staged input tile viewed through a
// descriptor whose start address advances by ky image rows = ky * 4096 B, a multiple of the 1024 B swizzle atom), the
// horizontal taps are folded into N, and the horizontal sum is seven warp shuffles per output channel in the epilogue
// (a thread owns the accumulator row of input pixel x; output pixel x needs rows x .. x+6 of the same warp).
// 28 MMAs per 4 x 26 output pixels instead of 196 per 128.
//
//  item      : 4 output rows x 26 output columns of one image (32 input columns, 10 input rows = ONE TMA box, 40 KB)
//  B operand : 7 tiles [32 (kx,o)][64 i] bf16, built once per CTA from the fp32 master weights (swizzled by hand)
//  roles     : warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue; persistent CTAs, TMEM double-buffered
#include "gconv.cuh"

namespace {

constexpr int C7_THREADS = 192;
constexpr int C7_STAGES = 3;
constexpr int C7_ROWS_IN = 10;                       // 4 output rows + 6
constexpr int C7_A_BYTES = C7_ROWS_IN * 32 * 128;    // 40 KB
constexpr int C7_B_BYTES = 7 * 32 * 128;             // 28 KB
constexpr int C7_SMEM = C7_STAGES * C7_A_BYTES + C7_B_BYTES + 1024 + 256;
constexpr int C7_TW = 26;                            // output columns per item

struct C7Dev {
  const float* w;          // master weights; element (o, ky, kx, i) at w_base + o*s_o + ky*s_ky + kx*s_kx + i*s_i
  long long w_base, s_o, s_ky, s_kx, s_i;
  const float* bias;       // [cout] or null
  bf16* out;
  long long o_str[3];      // x, y, n (elements)
  int cout;
  int hout, wout, n;
  int tiles_x, tiles_y;    // items per image
};

__global__ void __launch_bounds__(C7_THREADS)
    conv7few_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ C7Dev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C7_STAGES * C7_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + C7_B_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C7_STAGES;
  uint64_t* tmem_full = bars + 2 * C7_STAGES;          // [2]
  uint64_t* tmem_empty = tmem_full + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_img = p.tiles_x * p.tiles_y;
  const int nitems = per_img * p.n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    for (int s = 0; s < C7_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 4);                    // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<64>(tmem_slot);
  // B tiles: row r = kx*4 + o (rows 28..31 and o >= cout are zero), 64 input channels, K-major, 128-byte swizzle:
  // 16-byte chunk j of row r lives at chunk j ^ (r & 7) (tile bases are multiples of 1024 B)
  for (int e = threadIdx.x; e < 7 * 32 * 8; e += C7_THREADS) {
    const int j = e & 7, r = (e >> 3) & 31, ky = e >> 8;
    const int kx = r >> 2, o = r & 3;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (kx < 7 && o < p.cout) {
      float f[8];
      const float* src = p.w + p.w_base + o * p.s_o + ky * p.s_ky + kx * p.s_kx + (long long)(j * 8) * p.s_i;
#pragma unroll
      for (int t = 0; t < 8; ++t) f[t] = __ldg(src + t * p.s_i);
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(f[2 * t], f[2 * t + 1]);
    }
    *reinterpret_cast<uint4*>(sB + ky * 4096 + r * 128 + ((j ^ (r & 7)) << 4)) = u;
  }
  fence_proxy_async();                                 // generic-proxy writes -> visible to the MMA's async-proxy reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: one 32 x 10 pixel box per item =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int n = item / per_img, r = item - n * per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], C7_A_BYTES);
        tma_load_5d(sA + stage * C7_A_BYTES, &tmA, &full_bar[stage], 0, tx * C7_TW, ty * 4, 0, n);
        if (++stage == C7_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 32, 0, 0);
      const uint32_t b_base = smem_u32(sB);
      int stage = 0;
      uint32_t phase = 0;
      int li = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++li) {
        const int acc = li & 1;
        mbar_wait(&tmem_empty[acc], ((li >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + stage * C7_A_BYTES);
        const uint32_t d_addr = tmem_base + acc * 32;
#pragma unroll 1
        for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_desc_sw128(a_base + ky * 4096 + k * 32, 16, 1024);
            const uint64_t db = umma_desc_sw128(b_base + ky * 4096 + k * 32, 16, 1024);
            umma_bf16(d_addr, da, db, idesc, (ky | k) != 0);
          }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full[acc]);
        if (++stage == C7_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ================= epilogue: warp <-> TMEM lane quadrant warp % 4 = output row of the item =================
    const int q = warp & 3;
    float bias[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) bias[o] = (p.bias && o < p.cout) ? p.bias[o] : 0.f;
    int li = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++li) {
      const int n = item / per_img, r = item - n * per_img;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int acc = li & 1;
      mbar_wait(&tmem_full[acc], (li >> 1) & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 32), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      // out[x][o] = sum_kx D[x + kx][kx*4 + o]: lane x takes column kx*4+o from lane x + kx
      float s[4] = {bias[0], bias[1], bias[2], bias[3]};
#pragma unroll
      for (int kx = 0; kx < 7; ++kx)
#pragma unroll
        for (int o = 0; o < 4; ++o) s[o] += __shfl_down_sync(0xffffffffu, __uint_as_float(v[kx * 4 + o]), kx);
      const int oy = ty * 4 + q, ox = tx * C7_TW + lane;
      if (lane < C7_TW && oy < p.hout && ox < p.wout) {
        bf16* o = p.out + (long long)n * p.o_str[2] + (long long)oy * p.o_str[1] + (long long)ox * p.o_str[0];
        if (p.cout == 4) {
          uint2 u;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
          h[0] = __floats2bfloat162_rn(s[0], s[1]);
          h[1] = __floats2bfloat162_rn(s[2], s[3]);
          if ((reinterpret_cast<uintptr_t>(o) & 7) == 0) *reinterpret_cast<uint2*>(o) = u;
          else {
            o[0] = __float2bfloat16_rn(s[0]); o[1] = __float2bfloat16_rn(s[1]);
            o[2] = __float2bfloat16_rn(s[2]); o[3] = __float2bfloat16_rn(s[3]);
          }
        } else {
          for (int c = 0; c < p.cout; ++c) o[c] = __float2bfloat16_rn(s[c]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<64>(tmem_base);
  }
}

}  // namespace

// in: bf16 [N, hin, win, 64] (element strides in_str = x, y, n); out: bf16, (hin-6) x (win-6) x cout per image.
// w: fp32 master weights, element (o, ky, kx, i) at w[w_base + o*s_o + ky*s_ky + kx*s_kx + i*s_i].
extern "C" int dwc_conv7_few(const void* in, int n, int hin, int win, const int64_t* in_str, const float* w,
                             int64_t w_base, int64_t s_o, int64_t s_ky, int64_t s_kx, int64_t s_i, const float* bias,
                             int cout, void* out, const int64_t* out_str, dwc_stream_t stream) {
  DWC_CHECK(cout >= 1 && cout <= 4 && hin >= 7 && win >= 7 && n >= 1, "dwc_conv7_few: bad geometry");
  DWC_CHECK(in_str[0] == 64, "dwc_conv7_few: needs 64 contiguous input channels per pixel");
  CUtensorMap tmA;
  const int64_t dim[5] = {64, win, hin, 1, n};
  const int64_t str[5] = {1, in_str[0], in_str[1], in_str[2], in_str[2]};
  if (dwc_make_tmap5(&tmA, in, dim, str, 32, C7_ROWS_IN, 1, 1)) return 1;
  C7Dev d;
  d.w = w; d.w_base = w_base; d.s_o = s_o; d.s_ky = s_ky; d.s_kx = s_kx; d.s_i = s_i;
  d.bias = bias;
  d.out = reinterpret_cast<bf16*>(out);
  d.o_str[0] = out_str[0]; d.o_str[1] = out_str[1]; d.o_str[2] = out_str[2];
  d.cout = cout;
  d.hout = hin - 6; d.wout = win - 6; d.n = n;
  d.tiles_x = cdiv(d.wout, C7_TW);
  d.tiles_y = cdiv(d.hout, 4);
  static bool attr_set = false;
  if (!attr_set) {
    DWC_CUDA(cudaFuncSetAttribute(conv7few_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C7_SMEM));
    attr_set = true;
  }
  const long long nitems = (long long)d.tiles_x * d.tiles_y * n;
  int grid = dwc_num_sms();
  if (grid > nitems) grid = (int)nitems;
  conv7few_kernel<<<grid, C7_THREADS, C7_SMEM, as_stream(stream)>>>(tmA, d);
  DWC_LAUNCH_CHECK();
  return 0;
}
