// Shared between gconv.cu (tap-by-tap TMA kernel, SIMT kernel) and gconv_halo.cu (halo-tile kernel).
#pragma once
#include "common.cuh"

struct GConvDev {
  const void* a;
  const void* w;
  const float* bias;
  void* out;
  long long a_dim[5];
  long long a_str[5];
  long long o_str[3];
  int box_x, box_y, box_n;
  int tiles_x, tiles_y, tiles_n;
  int valid_x, valid_y, valid_n;
  int flat, flat_img, flat_pitch, flat_h, flat_w;
  int ntaps, C, K, ncols, ncols_padded;
  int out_dtype, accumulate;
  int nphase;            // >= 1
  long long phase_out_off;   // element offset of a phase's output; its weights start ncols_padded rows further down
  float2* stats;         // fused per-(n, tile, column) {sum, sum of squares} of the stored outputs, or null
  int debug;             // diagnostics only (DWC_GCONV_DEBUG): 1 no stores, 2 no epilogue, 3 no mainloop
  int taps[DWC_MAX_TAPS][3];
};

struct RowCoord {
  int x, y, n;
};

__device__ __forceinline__ RowCoord tile_row(const GConvDev& p, int tile, int r) {
  int tx = tile % p.tiles_x;
  int t2 = tile / p.tiles_x;
  int ty = t2 % p.tiles_y;
  int tn = t2 / p.tiles_y;
  RowCoord rc;
  rc.x = tx * p.box_x + r % p.box_x;
  int r2 = r / p.box_x;
  rc.y = ty * p.box_y + r2 % p.box_y;
  rc.n = tn * p.box_n + r2 / p.box_y;
  return rc;
}

// where (and whether) a row is stored
__device__ __forceinline__ bool out_offset(const GConvDev& p, const RowCoord& rc, long long* off) {
  if (p.flat) {
    int n = rc.x / p.flat_img;
    int rem = rc.x - n * p.flat_img;
    int yy = rem / p.flat_pitch;
    int xx = rem - yy * p.flat_pitch;
    if (n >= p.valid_n || yy >= p.flat_h || xx >= p.flat_w) return false;
    *off = (long long)n * p.o_str[2] + (long long)yy * p.o_str[1] + (long long)xx * p.o_str[0];
    return true;
  }
  if (rc.x >= p.valid_x || rc.y >= p.valid_y || rc.n >= p.valid_n) return false;
  *off = (long long)rc.n * p.o_str[2] + (long long)rc.y * p.o_str[1] + (long long)rc.x * p.o_str[0];
  return true;
}


// ---- fused statistics in the tcgen05 epilogues -------------------------------------------------------------------
// Column sums over the 32 lanes of a warp for 32 per-lane values (lane = accumulator row, index = column): after the
// five exchange stages lane j holds sum_l v_l[j] in v[0] - 31 shuffles instead of 5 per column.
template <int N> __device__ __forceinline__ void colsum_stage(float* v, int lane) {
  const bool up = (lane & N) != 0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float keep = up ? v[i + N] : v[i];
    const float send = up ? v[i] : v[i + N];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, N);
  }
}
__device__ __forceinline__ float warp_colsum32(float* v, int lane) {
  colsum_stage<16>(v, lane);
  colsum_stage<8>(v, lane);
  colsum_stage<4>(v, lane);
  colsum_stage<2>(v, lane);
  colsum_stage<1>(v, lane);
  return v[0];
}
// {sum, sum of squares} over the warp's 32 rows of the 32 columns f[0..32) of this lane's row, as they are STORED
// (rounded to bf16): lane j returns column j.  Rows outside the valid region contribute nothing.  The two reductions
// run one after the other so that only one 32-register scratch array is live next to f.
__device__ __forceinline__ float2 warp_col_stats32(const float* f, int lane, bool valid) {
  float t[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) t[j] = valid ? __bfloat162float(__float2bfloat16_rn(f[j])) : 0.f;
  const float s = warp_colsum32(t, lane);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float r = valid ? __bfloat162float(__float2bfloat16_rn(f[j])) : 0.f;
    t[j] = r * r;
  }
  return make_float2(s, warp_colsum32(t, lane));
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps

// CTA-pair (cta_group::2) variant of the tap-by-tap kernel, gconv2.cu; -1 = geometry not handled
int dwc_launch_gconv_tc2(const dwc_gconv_t* g, const GConvDev& d, cudaStream_t st);
// launches the halo-tile tcgen05 kernel (stride-1 k x k windows); defined in gconv_halo.cu
int dwc_launch_gconv_halo(const dwc_gconv_t* g, const GConvDev& d, int ksize, cudaStream_t st);
