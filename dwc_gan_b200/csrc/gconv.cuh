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
  int ksplit;            // one-wave kernel: CTAs of a cluster (1,1,ksplit) share a tile's K range (DSMEM reduction)
  void* out2;            // second, activated + reflect-haloed output (bf16, channels = ncols), or null
  int o2_halo, o2_layout, o2_act;
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


// ---- thread-block cluster helpers (split-K reduction through distributed shared memory) ----
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_map_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float ld_cluster_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// ---- second output of the tcgen05 epilogues: activation + reflect halo (+ parity planes) of the next block's input ----
// element offsets of the (up to 4) places interior pixel (x, y) of image n is written to in the haloed buffer: the pixel
// itself and its mirror images in the halo (x in [1, halo] <-> padded column halo - x; x in [W-1-halo, W-2] <-> padded
// column halo + 2(W-1) - x; same for rows).  Interior extent = valid_x x valid_y; at most one mirror per axis (H, W >= 4).
__device__ __forceinline__ long long out2_off(const GConvDev& p, int n, int Y, int X) {
  const int hp = p.valid_y + 2 * p.o2_halo, wp = p.valid_x + 2 * p.o2_halo;
  if (p.o2_layout == 0) return (((long long)n * hp + Y) * wp + X) * p.ncols;
  const int hq = hp >> 1, wq = wp >> 1;
  const int plane = ((Y & 1) << 1) | (X & 1);
  return ((((long long)n * 4 + plane) * hq + (Y >> 1)) * wq + (X >> 1)) * p.ncols;
}
__device__ __forceinline__ int out2_dests(const GConvDev& p, const RowCoord& rc, bool valid, long long* offs) {
  if (!valid) return 0;
  const int h = p.o2_halo, W = p.valid_x, H = p.valid_y;
  const int X0 = rc.x + h, Y0 = rc.y + h;
  int X1 = -1, Y1 = -1;
  if (h > 0) {
    if (rc.x >= 1 && rc.x <= h) X1 = h - rc.x;
    else if (rc.x >= W - 1 - h && rc.x <= W - 2) X1 = h + 2 * (W - 1) - rc.x;
    if (rc.y >= 1 && rc.y <= h) Y1 = h - rc.y;
    else if (rc.y >= H - 1 - h && rc.y <= H - 2) Y1 = h + 2 * (H - 1) - rc.y;
  }
  int nd = 0;
  offs[nd++] = out2_off(p, rc.n, Y0, X0);
  if (X1 >= 0) offs[nd++] = out2_off(p, rc.n, Y0, X1);
  if (Y1 >= 0) offs[nd++] = out2_off(p, rc.n, Y1, X0);
  if (X1 >= 0 && Y1 >= 0) offs[nd++] = out2_off(p, rc.n, Y1, X1);
  return nd;
}
// act(v) of 8 stored bf16 values, as the separate pass computes it (fp32 on the rounded value, rounded again)
__device__ __forceinline__ uint4 act8_bf16(const uint4& u, int act) {
  float f[8];
  Vec8<bf16>::load(reinterpret_cast<const bf16*>(&u), f);
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = act == 1 ? fmaxf(f[e], 0.f) : (f[e] > 0.f ? f[e] : 0.1f * f[e]);
  uint4 o;
  Vec8<bf16>::store(reinterpret_cast<bf16*>(&o), f);
  return o;
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
