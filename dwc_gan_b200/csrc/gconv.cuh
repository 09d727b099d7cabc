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


// CTA-pair (cta_group::2) variant of the tap-by-tap kernel, gconv2.cu; -1 = geometry not handled
int dwc_launch_gconv_tc2(const dwc_gconv_t* g, const GConvDev& d, cudaStream_t st);
// launches the halo-tile tcgen05 kernel (stride-1 k x k windows); defined in gconv_halo.cu
int dwc_launch_gconv_halo(const dwc_gconv_t* g, const GConvDev& d, int ksize, cudaStream_t st);
