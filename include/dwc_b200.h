/*
 * dwc_b200.h -- C ABI of the B200-native DWC-GAN training hot path.
 *
 * The reference (yhlleo/DWC-GAN) has no FFI layer: its hot path is stock torch.nn calls
 * (SURVEY.md 2.3).  This header is the net-new boundary those calls are replaced through
 * (level L2).  Each entry point cites the reference line(s) whose arithmetic
 * it takes over.  Conventions:
 *   - plain pointers and sizes only, no torch types; all pointers are DEVICE pointers
 *     unless a field says "host";
 *   - every call enqueues on `stream` and never synchronises or allocates;
 *   - return 0 on success, non-zero on error; dwc_last_error() gives the message
 *     (thread-local);
 *   - re-entrant and thread-safe (forward runs on the Python main thread, backward on
 *     autograd's device thread).
 *   - dtype codes: 0 = float32, 1 = bfloat16.  Activations are NHWC ("haloed buffers":
 *     [N, H+2h, W+2h, C], or four parity planes [N, 4, (H+2h)/2, (W+2h)/2, C] in front of
 *     a stride-2 convolution).
 */
#ifndef DWC_B200_H_
#define DWC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dwc_stream_t; /* cudaStream_t */

const char* dwc_last_error(void);
int dwc_abi_version(void);
/* 1 if the current device can run the tcgen05/TMA kernels (sm_100) and the driver exposes
 * cuTensorMapEncodeTiled. */
int dwc_tc_available(void);

#define DWC_F32 0
#define DWC_BF16 1
#define DWC_SIMT 0
#define DWC_TC 1
#define DWC_TC_HALO 2   /* gconv only: stride-1 k x k window, input staged once per 64-channel slab (gconv_halo.cu);
                         * box = (16,16,1): two 128-row accumulators per CTA, (8,16,1): one */
#define DWC_TC_HALO1 3  /* alias kept for diagnostics */
#define DWC_MAX_TAPS 64

/* ------------------------------------------------------------------------------------------
 * Generalised convolution as a multi-tap shifted GEMM ("gconv").
 *
 *   out[row, :] (+)= bias + sum_t  A[coord(row) + tap_t, 0:C] . W[:, t*C:(t+1)*C]^T
 *
 * A is a rank-5 element-strided view (C, X, Y, Z, N) of a haloed NHWC buffer, read with
 * zero fill outside [0, a_dim).  A tile is 128 rows, row r <-> (x, y, n) =
 * (x0 + r % bx, y0 + (r / bx) % by, n0 + r / (bx*by)).  With flat != 0, X is a flattened
 * (n, y, x) index over the input grid (image rows flat_img, pitch flat_pitch) and a row is
 * stored iff its decoded (y, x) lies inside flat_h x flat_w.
 *
 * Replaces: nn.Conv2d forward inside Conv2dBlock (networks/networks.py:531,577-580) and its
 * data gradient (aten::convolution_backward): the same kernel runs the
 * forward (reflect halo in A), the stride-1 dgrad (zero halo, flipped taps) and the four
 * parity phases of the stride-2 dgrad.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t dtype;        /* of A and W */
  int32_t backend;      /* DWC_SIMT or DWC_TC (TC needs bf16 and C % 64 == 0) */
  const void* a;
  int64_t a_dim[5];     /* C, X, Y, Z, N */
  int64_t a_str[5];     /* element strides, a_str[0] == 1 */
  int32_t box[3];       /* bx, by, bn ; bx*by*bn == 128 */
  int32_t tiles[3];     /* tiles along X, Y, N */
  int32_t valid[3];     /* rows with x >= valid[0] || y >= valid[1] || n >= valid[2] are dropped */
  int32_t flat, flat_img, flat_pitch, flat_h, flat_w;
  int32_t ntaps;
  const int32_t* taps;  /* HOST pointer: ntaps * 3 ints (dx, dy, z) */
  const void* w;        /* [ncols_padded][ntaps*C] */
  int32_t ncols, ncols_padded;
  const float* bias;    /* [ncols] or NULL */
  void* out;
  int64_t o_str[3];     /* element strides of out for x, y, n (columns contiguous) */
  int32_t out_dtype;
  int32_t accumulate;   /* out += ... */
  /* Phases: nphase (0 or 1 = single) independent problems that share A, the tiling and the taps; phase ph uses the
   * weight matrix at w + ph*phase_w_off elements and writes to out + ph*phase_out_off elements.  The four parity
   * phases of a stride-2 data gradient are one launch this way. */
  int32_t nphase;
  int32_t reserved0;
  int64_t phase_w_off;
  int64_t phase_out_off;
  /* Optional fused statistics (tcgen05 backend, bf16 output, box[2] == 1, no flat / phases, ncols % 64 == 0): the
   * epilogue also writes per-(n, tile, column) partial sums of the stored (bf16-rounded) outputs,
   * stats[((n * splits + t) * ncols + c) * 2 + {0, 1}] = {sum, sum of squares} with splits = tiles[0] * tiles[1] and
   * t = ty * tiles[0] + tx - the first half of InstanceNorm / AdaIN / LayerNorm (networks.py:545,706-719,736-752)
   * without re-reading the convolution's output (same layout as dwc_nc_stats).  NULL = off. */
  float* stats;
  /* Optional second, ACTIVATED output (same restrictions as `stats`, plus ncols a multiple of the column tile): the
   * epilogue also writes act(stored value) - act 1 ReLU, 2 LeakyReLU(0.1) - into the haloed buffer `out2` of interior
   * extent valid[1] x valid[0], `out2_halo` reflect-halo pixels (written from the interior pixel they mirror) and layout
   * 0 / 1 (dwc_hbuf_t conventions, channels = ncols): activation + nn.ReflectionPad2d of the NEXT Conv2dBlock for the
   * norm-less layers (style encoder, discriminator: networks.py:531,556-567) without a separate pass.  NULL = off. */
  void* out2;
  int32_t out2_halo, out2_layout, out2_act, reserved1;
} dwc_gconv_t;

int dwc_gconv(const dwc_gconv_t* p, dwc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Weight gradient: dw[ia*s_a + t*s_t + ib*s_b] (+)= sum_pixels A[pix, ia] * B[pix + tap_t, ib]
 * and optionally dbias[ia] += sum_pixels A[pix, ia].   A = dY view, B = padded input view,
 * both rank-5 (C, X, Y, Z, N); pixel tiles enumerated exactly as in gconv.
 * Replaces the weight/bias part of aten::convolution_backward.
 * Split-K partial sums go through `workspace` and are reduced in a fixed order
 * (deterministic).  dwc_wgrad_workspace_bytes() returns the size to provide.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t dtype, backend;
  const void* a; int64_t a_dim[5]; int64_t a_str[5];
  const void* b; int64_t b_dim[5]; int64_t b_str[5];
  int32_t box[3];
  int32_t tiles[3];
  int32_t ntaps;
  const int32_t* taps;  /* HOST pointer, (dx, dy, z) applied to B */
  int32_t ca, cb;       /* channels of A and of B */
  float* dw; int64_t s_a, s_t, s_b;
  float* dbias;         /* NULL to skip */
  int32_t accumulate;   /* dw/dbias += (else overwritten) */
  float* workspace; int64_t workspace_bytes;
  /* Optional index split for row-im2col operands (8-pixel x 8-channel windows of a few-channel image): with
   * remap_axis = 1 (A index) or 2 (B index) the channel index i of that operand is stored at
   * (i / remap_div) * remap_hi_stride + (i % remap_div) * remap_lo_stride instead of i * s_a (or s_b), and dropped
   * unless i % remap_div < remap_lo_limit and i / remap_div < remap_hi_limit.  remap_axis = 0: plain strides. */
  int32_t remap_axis, remap_div, remap_lo_limit, remap_hi_limit;
  int64_t remap_hi_stride, remap_lo_stride;
} dwc_wgrad_t;

int64_t dwc_wgrad_workspace_bytes(const dwc_wgrad_t* p);
int dwc_wgrad(const dwc_wgrad_t* p, dwc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Haloed-buffer geometry shared by the element-wise kernels.
 * layout 0: [N, H+2*halo, W+2*halo, C]; layout 1: parity planes [N, 4, (H+2*halo)/2, ...].
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  void* ptr;            /* start of the allocation (not the interior) */
  int32_t n, h, w, c;   /* interior extent */
  int32_t halo;
  int32_t layout;
  int32_t dtype;
} dwc_hbuf_t;

/* per-(n,c) sum and sum of squares over the interior, in `splits` pixel slices that the finalize
 * kernels add up in a fixed order: stats[(n*splits+s)*C+c] = {sum, sumsq} (float2).
 * First half of nn.InstanceNorm2d / AdaIN / LayerNorm statistics (networks.py:545,706-719,736-752). */
int dwc_nc_stats(const dwc_hbuf_t* y, int splits, float* stats, dwc_stream_t stream);

/* Turn statistics into per-(n,c) scale/shift.  kind: 0 none (scale 1, shift 0), 1 instance norm,
 * 2 AdaIN (weight,bias [N,C]), 3 MUNIT LayerNorm (gamma,beta [C]; unbiased std, eps added to std).
 * coef[n*C+c] = {scale, shift, mean, rstd}   (float4) */
int dwc_norm_finalize(int kind, const float* stats, int splits, int n, int c, int hw, float eps,
                      const float* weight, const float* bias, float* coef, dwc_stream_t stream);

/* out = reflect_pad( act(scale*y + shift) + residual ).  act: 0 none 1 relu 2 lrelu(0.1).
 * `res` may be NULL.  The whole haloed extent of `out` is written.
 * Replaces norm + activation + residual add + nn.ReflectionPad2d of the next Conv2dBlock
 * (networks.py:514-522,531,580-585). */
int dwc_post_fwd(const dwc_hbuf_t* y, const float* coef, int act, const dwc_hbuf_t* res,
                 const dwc_hbuf_t* out, dwc_stream_t stream);

/* Backward, pass 1: red[(n*splits+s)*C+c] = {sum dz, sum dz*y} with dz = fold(dout) * act'(scale*y+shift). */
/* 1 if dwc_post_bwd_reduce(..., prefolded = 2, ...) can fold dout's reflect-halo gradient into the interior itself, in place
 * and bit-identically to dwc_fold_halo, while it streams dout for the reduction (one launch less per norm site). */
int dwc_post_bwd_reduce_can_fold(const dwc_hbuf_t* dout, const dwc_hbuf_t* y);   /* geometry allows it */
int dwc_post_bwd_reduce_folds(const dwc_hbuf_t* dout, const dwc_hbuf_t* y);      /* ... and DWC_FOLD_IN_REDUCE=1 (measured slower: off) */
int dwc_post_bwd_reduce(const dwc_hbuf_t* dout, const dwc_hbuf_t* y, const float* coef, int act, int splits,
                        float* red, int prefolded, dwc_stream_t stream);
/* Backward, pass 1b: per-(n,c) coefficients bco = {a, b, c, 0} so that dy = a*dz + b*y + c,
 * plus parameter gradients: AdaIN dweight/dbias [N,C] (overwritten), LayerNorm dgamma/dbeta [C]
 * (accumulated). */
int dwc_norm_bwd_finalize(int kind, const float* red, int splits, const float* coef, int n, int c, int hw, float eps,
                          const float* weight, float* dweight, float* dbias, float* bco,
                          dwc_stream_t stream);
/* Backward, pass 2: dy = a*dz + b*y + c written with a ZERO halo; dres (optional, same geometry
 * as the forward `res`) = fold(dout) in the interior and zero in the halo. */
int dwc_post_bwd_apply(const dwc_hbuf_t* dout, const dwc_hbuf_t* y, const float* coef, const float* bco,
                       int act, const dwc_hbuf_t* dy, const dwc_hbuf_t* dres, int prefolded, dwc_stream_t stream);

/* dwc_norm_finalize + dwc_post_fwd in one call: on the row-streaming path (bf16, rows of >= 2 KB, C <= 512) the
 * coefficients are computed inside the normalise kernel from the partial statistics (each CTA recomputes its sample's
 * C coefficients from L2) and `coef` [N,C,4] is written for the backward pass; otherwise the two passes are launched. */
int dwc_post_fwd_norm(const dwc_hbuf_t* y, int kind, const float* stats, int splits, float eps, const float* weight,
                      const float* bias, int act, const dwc_hbuf_t* res, const dwc_hbuf_t* out, float* coef,
                      dwc_stream_t stream);
/* dwc_norm_bwd_finalize + dwc_post_bwd_apply in one call (same rule; `bco` [N,C,4] is scratch for the two-pass path). */
int dwc_post_bwd_apply_norm(const dwc_hbuf_t* dout, const dwc_hbuf_t* y, const float* coef, int kind, const float* red,
                            int splits, float eps, const float* weight, float* dweight, float* dbias, float* bco,
                            int act, const dwc_hbuf_t* dy, const dwc_hbuf_t* dres, int prefolded, dwc_stream_t stream);

/* nn.Upsample(2, bilinear, align_corners=False) + reflect pad (networks_v2.py:154, networks.py:531). */
int dwc_upsample_pad_fwd(const dwc_hbuf_t* x, const dwc_hbuf_t* out, dwc_stream_t stream);
int dwc_upsample_pad_bwd(const dwc_hbuf_t* dout, const dwc_hbuf_t* dx, int prefolded, dwc_stream_t stream);
/* Fused InstanceNorm (kind 1) / AdaIN (kind 2) site for small feature maps (dwc_post_fused_ok: bf16, H*W <= 1536,
 * C % 32 == 0): one kernel per site and direction, each (sample, channel group) slab staged once in shared memory.
 * Forward = dwc_nc_stats + dwc_norm_finalize + dwc_post_fwd (also writes coef [N,C,4] for the backward);
 * backward = dwc_fold_halo + dwc_post_bwd_reduce + dwc_norm_bwd_finalize + dwc_post_bwd_apply (dout is NOT modified).
 * Same arithmetic as the separate passes (networks.py:514-522,545,693-722). */
int dwc_post_fused_ok(int c, int hw, int dtype);
int dwc_post_fused_fwd(const dwc_hbuf_t* y, int kind, const float* weight, const float* bias, float eps, int act,
                       const dwc_hbuf_t* res, const dwc_hbuf_t* out, float* coef, dwc_stream_t stream);
int dwc_post_fused_bwd(const dwc_hbuf_t* dout, const dwc_hbuf_t* y, const float* coef, int kind, int act,
                       const float* weight, float* dweight, float* dbias, const dwc_hbuf_t* dy,
                       const dwc_hbuf_t* dres, dwc_stream_t stream);
/* In-place backward of a reflect halo: every interior pixel whose mirror images lie in the halo receives their sum
 * (the halo itself is left as is).  After it the three backward passes above take prefolded = 1 and stream the
 * interior without gathering reflections.  Needs h, w >= 2*halo + 2. */
int dwc_fold_halo(const dwc_hbuf_t* d, dwc_stream_t stream);

/* NCHW float32 image -> haloed NHWC buffer, optional 2x2 average pooling first
 * (F.interpolate(0.5, bilinear) == avg_pool2d, networks.py:113) and reflect padding. */
int dwc_image_pad_fwd(const float* img, int n, int c, int h, int w, int pool, const dwc_hbuf_t* out,
                      dwc_stream_t stream);
int dwc_image_pad_bwd(const dwc_hbuf_t* dout, int pool, float* dimg, int n, int c, int h, int w,
                      int accumulate, dwc_stream_t stream);

/* Row-im2col of a few-channel image for the tensor-core path of the first convolutions (3 -> 64, 7x7 s1 and
 * 4x4 s2): rows[n, z, yr, x, j*8 + ch] = P[n, yr*ys + z, x*sx + j, ch] for j, ch < 8 where P is the (avg-pooled,)
 * reflect-padded image; ys = 2 stores even/odd padded rows as two planes z.  0 beyond the image. */
int dwc_image_rows_fwd(const float* img, int n, int c, int h, int w, int pool, int pad, int sx, int ys, int wo,
                       void* rows, int dtype, dwc_stream_t stream);
/* Backward of the fused decoder heads for the tensor-core path: from dimg/datt (may be NULL) and the saved
 * img/att builds  rows_d[n, Y, X, j*8+c] = dyz[n, Y, X+j, c]   (dyz = dy with a zero halo of `halo`, for dgrad),
 *                 win[n, h, u, j*8+c]   = dy[n, h, u-j, c]     (u over the padded width w+halo, for wgrad)
 * and per-block partial sums of dy (bias gradient), part[block*4 + c]; returns the number of blocks in *nblocks. */
int dwc_heads_bwd_rows(const float* dimg, const float* datt, const float* img, const float* att, int n, int h, int w,
                       int halo, void* rows_d, void* win, int dtype, float* part, int32_t* nblocks, dwc_stream_t stream);

/* Decoder heads: y[..., 0:3] -> tanh -> img NCHW f32 ; y[..., 3] -> sigmoid -> att NCHW f32
 * (networks_v2.py:162-169).  Backward writes dy (zero halo). */
int dwc_heads_fwd(const dwc_hbuf_t* y, float* img, float* att, dwc_stream_t stream);
int dwc_heads_bwd(const float* dimg, const float* datt, const float* img, const float* att,
                  const dwc_hbuf_t* dy, dwc_stream_t stream);

/* x = img*att + real*(1-att)   (solver.py:160-161).  Backward: dimg, datt. */
int dwc_blend_fwd(const float* img, const float* att, const float* real, float* out, int n, int c, int hw,
                  dwc_stream_t stream);
int dwc_blend_bwd(const float* dout, const float* img, const float* att, const float* real, float* dimg,
                  float* datt, int n, int c, int hw, dwc_stream_t stream);

/* Global average pool of relu(y) (StyleEncoder tail: ReLU + AdaptiveAvgPool2d(1), networks_v2.py:113). */
int dwc_relu_gap_fwd(const dwc_hbuf_t* y, float* out, dwc_stream_t stream);
int dwc_relu_gap_bwd(const float* dout, const dwc_hbuf_t* y, const dwc_hbuf_t* dy, dwc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Small dense algebra (style MLP, mapping network, heads, text encoder): fp32 accumulate.
 *   C[m,n] = act( alpha * sum_k A[m,k]*B[k,n] + bias[n] + beta*C[m,n] )
 * with arbitrary element strides.  a_dtype applies to A only (B, C, bias are float32).
 * Replaces nn.Linear stacks (networks.py:496-499, networks_v2.py:117-127,208-210).
 * ------------------------------------------------------------------------------------------ */
int dwc_sgemm(int m, int n, int k, float alpha, const void* a, int a_dtype, int64_t a_sm, int64_t a_sk,
              const float* b, int64_t b_sk, int64_t b_sn, float beta, float* c, int64_t c_sm, int64_t c_sn,
              const float* bias, int act, dwc_stream_t stream);

/* The same GEMM on the tensor cores: tcgen05.mma kind::tf32 with fp32 operands fetched by TMA (no conversion pass),
 * fp32 accumulation (csrc/dense_tc.cu).  A K-major (a_sk == 1) or M-major (a_sm == 1), B K-major (b_sk == 1) or
 * N-major (b_sn == 1), C row-major; leading strides multiples of 4 elements, bases 16-byte aligned
 * (dwc_gemm_tf32_ok).  dwc_set_tf32(1) makes dwc_sgemm / dwc_sgemm_ws route eligible problems here (bf16 product
 * mode); with 0 (default, fp32 validation mode) they stay exact fp32. */
int dwc_gemm_tf32_ok(int m, int n, int k, const void* a, int64_t a_sm, int64_t a_sk, const void* b, int64_t b_sk,
                     int64_t b_sn, const void* c, int64_t c_sm, int64_t c_sn);
int dwc_gemm_tf32(int m, int n, int k, float alpha, const float* a, int64_t a_sm, int64_t a_sk, const float* b,
                  int64_t b_sk, int64_t b_sn, float beta, float* c, int64_t c_sm, const float* bias, int act,
                  dwc_stream_t stream);
void dwc_set_tf32(int on);
int dwc_get_tf32(void);
/* Same with a caller-provided float workspace of dwc_sgemm_workspace_bytes(m, n, k) bytes: problems with few
 * output tiles and a long K (head layers at small batch) are then split along K over many CTAs and summed in a
 * fixed order (deterministic). */
int64_t dwc_sgemm_workspace_bytes(int m, int n, int k);
int dwc_sgemm_ws(int m, int n, int k, float alpha, const void* a, int a_dtype, int64_t a_sm, int64_t a_sk,
                 const float* b, int64_t b_sk, int64_t b_sn, float beta, float* c, int64_t c_sm, int64_t c_sn,
                 const float* bias, int act, float* workspace, int64_t workspace_bytes, dwc_stream_t stream);
/* out[n] (+)= sum_m A[m,n]  (bias gradients) */
int dwc_colsum(int m, int n, const float* a, int64_t a_sm, int64_t a_sn, float* out, int accumulate,
               dwc_stream_t stream);
/* element-wise helpers on float32: relu backward from output, dropout-mask multiply */
int dwc_relu_bwd(const float* dout, const float* out, float* din, int64_t count, dwc_stream_t stream);
int dwc_mul(const float* a, const float* b, float* out, int64_t count, dwc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Text encoder (networks_v2.py:213-254): embedding + style concat, bidirectional LSTM layers.
 * ------------------------------------------------------------------------------------------ */
int dwc_embed_concat_fwd(const int64_t* tokens /*[B,T]*/, const float* emb, const float* style, const float* mask,
                         float* x /*[T,B,E+S]*/, int b, int t, int e, int s, dwc_stream_t stream);
int dwc_embed_concat_bwd(const int64_t* tokens, const float* dx, const float* mask, float* demb, float* dstyle,
                         int b, int t, int e, int s, int pad_idx, dwc_stream_t stream);
/* Whole-layer bidirectional LSTM recurrence as ONE persistent cooperative kernel (csrc/lstm.cu): both
 * directions, all time steps up to max(lens); gate order i,f,g,o; per-sample lengths give
 * pack_padded_sequence semantics (state frozen and output zero at t >= lens[b]).
 *   xproj [T,B,2,4H]  input projection + both biases;  whh [2,4H,H] recurrent weights (nn.LSTM layout,
 *   forward then reverse);  out [T,B,2H] (may be NULL);  gates_save [T,B,2,4H] (activated gates) and
 *   c_save [T,B,2,H] (cell state after the step) for the backward pass (both NULL in inference);
 *   h_final / c_final [2,B,H];  workspace: dwc_lstm_workspace_bytes(b, h) bytes.
 * Replaces nn.LSTM forward (networks_v2.py:225-233) after the input-projection GEMM. */
int64_t dwc_lstm_workspace_bytes(int b, int h);
int dwc_lstm_layer_fwd(int t_total, int b, int h, const float* xproj, const float* whh, const int64_t* lens,
                       float* out, float* gates_save, float* c_save, float* h_final, float* c_final,
                       void* workspace, dwc_stream_t stream);
/* Backward of the recurrence: dout [T,B,2H] (gradient of `out`, may be NULL), dh_final / dc_final [2,B,H]
 * (gradients of the final states) -> dgates [T,B,2,4H] (gradient of the gate pre-activations, zero at
 * padded steps), from which the caller forms the weight / input gradients with GEMMs. */
int dwc_lstm_layer_bwd(int t_total, int b, int h, const float* whh, const int64_t* lens, const float* dout,
                       const float* gates_save, const float* c_save, const float* dh_final,
                       const float* dc_final, float* dgates, void* workspace, dwc_stream_t stream);

/* dst[b][c][r] = src[b][r][c] (float32) */
int dwc_transpose(const float* src, float* dst, int batch, int rows, int cols, dwc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * GMM style space (tools.py:65-70, gmm.py:13-22) and losses.
 * ------------------------------------------------------------------------------------------ */
/* z[b, j*cdim+k] = mu[b,j] + stddev * eps[k, b, j]   (eps laid out (1, cdim, B, ncls)) */
int dwc_gmm_sample(const float* mu, const float* eps, float stddev, float* z, int b, int ncls, int cdim,
                   dwc_stream_t stream);
/* loss = sum_i mean_b sum_k 0.5*(log(sigma/exp(lv)) + (exp(lv)+(mu-c[b,i])^2)/sigma - 1);
 * mu, lv [B, ncls*cdim]; also writes dmu, dlv (gradient of the loss, unscaled). */
int dwc_gmm_kl(const float* mu, const float* lv, const float* c, float sigma, float* loss, float* dmu,
               float* dlv, int b, int ncls, int cdim, dwc_stream_t stream);
/* loss[0] = mean |a-b| (a,b flat of a_dtype/b_dtype)  (solver.py:113-114,127-132).  `loss` is a ZEROED scratch of
 * DWC_L1_SCRATCH floats: [0] the result, the rest a ticket and per-block partial sums added in a fixed order, so the
 * value is bit-reproducible.
 * Backward: da = gscale[0]*sign(a-b)/count, db = -da; either may be NULL; gscale is a device scalar. */
#define DWC_L1_SCRATCH 600
int dwc_l1_loss_fwd(const void* a, int a_dtype, const void* b, int b_dtype, int64_t count, float* loss,
                    dwc_stream_t stream);
int dwc_l1_loss_bwd(const void* a, int a_dtype, const void* b, int b_dtype, int64_t count, const float* gscale,
                    void* da, void* db, dwc_stream_t stream);
/* loss[0] = mean (x - target)^2 ; dx = gscale*2(x-target)/count  (LSGAN, networks.py:131,158) */
int dwc_mse_const_loss_fwd(const float* x, float target, int64_t count, float* loss, dwc_stream_t stream);
int dwc_mse_const_loss_bwd(const float* x, float target, int64_t count, const float* gscale, float* dx,
                           dwc_stream_t stream);
/* loss[0] = mean BCE-with-logits(x, y); dx = gscale*(sigmoid(x)-y)/count (networks.py:83) */
int dwc_bce_logits_loss_fwd(const float* x, const float* y, int64_t count, float* loss, dwc_stream_t stream);
int dwc_bce_logits_loss_bwd(const float* x, const float* y, int64_t count, const float* gscale, float* dx,
                            dwc_stream_t stream);

/* One discriminator scale's adversarial terms fused (MsImageDis.calc_dis_loss / calc_gen_loss, networks.py:116-170, as
 * the Solver batches them, solver.py:206-207,333-334): loss[0] = sum_k weight_k * mean over rows [row0,row1) of
 *   kind 0: (src - target)^2          src [src_rows, src_cols] fp32 (the 1-channel patch output)
 *   kind 1: BCE-with-logits(cls, y)   cls [cls_rows, cls_cols] fp32, y = labels [row1-row0, cls_cols] (same for all terms)
 * Backward writes the FULL gradients dsrc / dcls (zero outside every term's rows), scaled by the device scalar gscale. */
#define DWC_ADV_MAX_TERMS 8
typedef struct {
  int32_t kind, row0, row1;
  float target, weight;
} dwc_adv_term_t;
int dwc_adv_loss_fwd(const float* src, int src_rows, int src_cols, const float* cls, int cls_rows, int cls_cols,
                     const float* labels, const dwc_adv_term_t* terms, int nterms, float* loss, dwc_stream_t stream);
int dwc_adv_loss_bwd(const float* src, int src_rows, int src_cols, const float* cls, int cls_rows, int cls_cols,
                     const float* labels, const dwc_adv_term_t* terms, int nterms, const float* gscale, float* dsrc,
                     float* dcls, dwc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer: torch.optim.Adam with coupled L2 (solver.py:65-68) + EMA (utils.py:52-54) on flat
 * buffers, chunked.  chunk table entries: {offset, length, active}.  hyper (device, float[8]):
 * {lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2, grad_scale}.
 * ------------------------------------------------------------------------------------------ */
int dwc_adam_step(float* param, const float* grad, float* m, float* v, int64_t count,
                  const uint8_t* active /*per 1024-element chunk, device, may be NULL*/,
                  const float* hyper, dwc_stream_t stream);
int dwc_ema_step(const float* param, float* avg, int64_t count, float beta, dwc_stream_t stream);
/* 7x7 stride-1 valid convolution from 64 channels to cout <= 4 channels (bf16 in / out, fp32 accumulation) on tcgen05:
 * the decoder heads (networks_v2.py:162-169, reflect-haloed input) and the image gradient of a first encoder
 * convolution (transpose of networks_v2.py:52, zero-haloed output gradient, flipped taps).  Vertical taps are folded
 * into K, horizontal taps into N (7 x 4 columns) and summed by warp shuffles in the epilogue.
 * in: [n, hin, win, 64], element strides in_str = {x, y, n}; out: (hin-6) x (win-6) x cout per image, out_str = {x, y, n}.
 * The weights are read from the fp32 master copy: element (o, ky, kx, i) at w[w_base + o*s_o + ky*s_ky + kx*s_kx + i*s_i]
 * (strides may be negative: a data gradient walks the taps backwards).  bias may be NULL. */
int dwc_conv7_few(const void* in, int n, int hin, int win, const int64_t* in_str, const float* w, int64_t w_base,
                  int64_t s_o, int64_t s_ky, int64_t s_kx, int64_t s_i, const float* bias, int cout, void* out,
                  const int64_t* out_str, dwc_stream_t stream);

/* master float32 weights [Cout][taps][Cin] -> packed compute-dtype GEMM operands.
 * mode 0: forward  wf[co][t][ci]                (cast only, rows padded to ncols_padded)
 * mode 1: stride-1 dgrad  wd[ci][T-1-t][co]     (taps reversed)
 * mode 2: stride-2 k4 dgrad, 4 parity phases  wd[phase][ci][(i,j)][co]
 * mode 3: forward over a row-im2col input     wr[co][kh][j*8+ci] = w[co][kh][j][ci]  (j < kw, ci < cin, else 0)
 * mode 4: stride-1 dgrad over row-im2col dY   wr[ci][kh'][j*8+co] = w[co][KH-1-kh'][KW-1-j][ci]  (else 0) */
int dwc_pack_weights(const float* w, int cout, int taps_h, int taps_w, int cin, int mode, void* out,
                     int out_dtype, int rows_padded, dwc_stream_t stream);
/* The same for every operand of a network in one launch: `table_dev` is a DEVICE array of `count` entries (built once
 * by the host; the pointers it holds - slices of the flat parameter buffer and the in-place rewritten operands - do not
 * change between optimizer steps).  total = number of elements of `out`. */
typedef struct {
  const float* w;
  void* out;
  int32_t cout, kh, kw, cin, mode, rows_padded, out_dtype, reserved;
  int64_t total;
} dwc_pack_entry_t;
int dwc_pack_weights_batch(const dwc_pack_entry_t* table_dev, int count, dwc_stream_t stream);
int dwc_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t count, dwc_stream_t stream);
int dwc_fill(void* dst, int dtype, float value, int64_t count, dwc_stream_t stream);

#ifdef DWC_EXPERIMENTAL   /* parked kernels, csrc/experimental/ (not in the default library) */
/* One-pass backward of an InstanceNorm (kind 1) / AdaIN (kind 2) site with a whole sample held in the shared memory of
 * a thread-block cluster (8 CTAs, partial sums exchanged through distributed shared memory): reflect-halo fold of
 * `dout` (halo <= 1, not modified), per-channel reductions, coefficients, dy (zero halo), optional residual gradient
 * `dres` and the AdaIN parameter gradients in ONE launch and one read of dout and y.  Same arithmetic as
 * dwc_fold_halo + dwc_post_bwd_reduce + dwc_post_bwd_apply_norm (networks.py:545,693-722 backward).
 * dwc_post_bwd_cluster_ok() says whether a site qualifies (bf16, plain layouts, H %% 8 == 0, rows fit shared memory:
 * the 256-channel 32x32 residual blocks). */
int dwc_post_bwd_cluster_ok(const dwc_hbuf_t* dout, const dwc_hbuf_t* y, int kind, const dwc_hbuf_t* dy,
                            const dwc_hbuf_t* dres);
int dwc_post_bwd_cluster(const dwc_hbuf_t* dout, const dwc_hbuf_t* y, const float* coef, int kind, int act,
                         const float* weight, float* dweight, float* dbias, const dwc_hbuf_t* dy,
                         const dwc_hbuf_t* dres, dwc_stream_t stream);
#endif

#ifdef __cplusplus
}
#endif
#endif /* DWC_B200_H_ */
