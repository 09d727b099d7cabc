// Dense fp32 GEMMs of the text encoder / style MLP on the tensor cores: tcgen05.mma kind::tf32, operands fetched by TMA
// straight from the fp32 tensors (no conversion pass), fp32 accumulation in TMEM.
//
//   C[M,N] = alpha * A[M,K] * B[K,N] + beta * C + bias[n]   (optional ReLU)
//
// A may be K-major (a_sk == 1: activations [rows, features]) or M-major (a_sm == 1: a transposed view, as in
// dW = dY^T X); B may be K-major (b_sk == 1: an nn.Linear weight [N, K]) or N-major (b_sn == 1: dX = dY W).  Used for the
// LSTM input projections over all packed tokens (M = T*B, N = 2400: networks_v2.py:197-203,225-233), their data / weight
// gradients, and the larger nn.Linear layers (networks.py:496-499, networks_v2.py:117-127,208-210) in the bf16 product
// mode; the fp32 validation mode keeps the exact fp32 SIMT kernels of dense.cu.  TF32 keeps 10 mantissa bits of each
// operand: ~3e-4 relative on a dot product of a few hundred terms.
//
// One CTA per 128 x BN output tile (BN = 64 / 128 / 256), K in slabs of 32 fp32 (= one 128-byte swizzle row):
// warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (TMEM -> registers -> global).
#include "common.cuh"
#include <cudaTypedefs.h>

namespace {

constexpr int GT_BM = 128;
constexpr int GT_BK = 32;               // fp32 elements per slab = 128 bytes
constexpr int GT_THREADS = 192;
constexpr int GT_ATOM = 32 * 128;       // MN-major atom: 32 k-rows x 32 elements = 4 KB

struct GemmTcP {
  int M, N, K;
  int a_mn, b_mn;                       // operand is MN-major (else K-major)
  float* C;
  long long c_sm;
  const float* bias;
  float alpha, beta;
  int act;
};

template <int BN> struct GtCfg {
  static constexpr int A_BYTES = GT_BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN >= 256 ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int SMEM = STAGES * STAGE + 1024 + 256;
};

// instruction descriptor, kind::tf32: c_format F32 (1) at [4,6), a/b format TF32 (2) at [7,10) / [10,13)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// UMMA shared-memory descriptor with an explicit layout type: 2 = SWIZZLE_128B (K-major operands), 1 =
// SWIZZLE_128B_BASE32B - the 128-byte swizzle on 32-byte chunks over 4-row atoms that MN-major 32-bit operands need (the
// tensor core transposes at element granularity; TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
__device__ __forceinline__ uint64_t umma_desc_lt(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

template <int BN>
__global__ void __launch_bounds__(GT_THREADS)
    gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ GemmTcP p) {
  using Cfg = GtCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * GT_BM, n0 = blockIdx.y * BN;
  const int nslab = (p.K + GT_BK - 1) / GT_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nslab; ++kb) {
        const int k0 = kb * GT_BK;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], Cfg::STAGE);        // out-of-bounds parts of a box are zero-filled and counted
        uint8_t* sa = smem + stage * Cfg::STAGE;
        uint8_t* sb = sa + Cfg::A_BYTES;
        if (p.a_mn) {
#pragma unroll
          for (int j = 0; j < GT_BM / 32; ++j) tma_load_2d(sa + j * GT_ATOM, &tmA, &full_bar[stage], m0 + 32 * j, k0);
        } else {
          tma_load_2d(sa, &tmA, &full_bar[stage], k0, m0);
        }
        if (p.b_mn) {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) tma_load_2d(sb + j * GT_ATOM, &tmB, &full_bar[stage], n0 + 32 * j, k0);
        } else {
          tma_load_2d(sb, &tmB, &full_bar[stage], k0, n0);
        }
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(GT_BM, BN, p.a_mn, p.b_mn);
      // K-major (SWIZZLE_128B): 8-element k steps are 32 bytes apart inside the swizzled rows, 8-row groups SBO = 1024 B;
      // MN-major (SWIZZLE_128B_BASE32B): 8 k-rows = 1024 B per step = two 4-row swizzle atoms SBO = 512 B apart,
      // 32-element atoms along M / N are LBO = 4 KB apart
      const uint32_t a_step = p.a_mn ? 1024u : 32u, b_step = p.b_mn ? 1024u : 32u;
      const uint32_t a_lbo = p.a_mn ? (uint32_t)GT_ATOM : 16u, b_lbo = p.b_mn ? (uint32_t)GT_ATOM : 16u;
      const uint32_t a_sbo = p.a_mn ? 512u : 1024u, b_sbo = p.b_mn ? 512u : 1024u;
      const uint32_t a_lt = p.a_mn ? 1u : 2u, b_lt = p.b_mn ? 1u : 2u;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nslab; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE);
        const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < GT_BK / 8; ++k) {
          const uint64_t da = umma_desc_lt(a_addr + k * a_step, a_lbo, a_sbo, a_lt);
          const uint64_t db = umma_desc_lt(b_addr + k * b_step, b_lbo, b_sbo, b_lt);
          umma_tf32(tmem_base, da, db, idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tmem_full);
    }
  } else {
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int cc = 0; cc < BN; cc += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
      tmem_ld_wait();
      const int c0 = n0 + cc;
      if (row >= p.M || c0 >= p.N) continue;
      float* o = p.C + (long long)row * p.c_sm + c0;
      const bool vec = c0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float f[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          f[e] = p.alpha * __uint_as_float(v[j + e]);
          if (p.bias && c0 + j + e < p.N) f[e] += __ldg(p.bias + c0 + j + e);
        }
        if (vec) {
          if (p.beta != 0.f) {
            const float4 old = *reinterpret_cast<const float4*>(o + j);
            f[0] += p.beta * old.x; f[1] += p.beta * old.y; f[2] += p.beta * old.z; f[3] += p.beta * old.w;
          }
          if (p.act == 1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) f[e] = fmaxf(f[e], 0.f);
          }
          *reinterpret_cast<float4*>(o + j) = make_float4(f[0], f[1], f[2], f[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (c0 + j + e < p.N) {
              float r = f[e] + (p.beta != 0.f ? p.beta * o[j + e] : 0.f);
              if (p.act == 1) r = fmaxf(r, 0.f);
              o[j + e] = r;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BN>(tmem_base);
  }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_tiled_fn get_encode_f32() {
  static encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(ptr);
  }
  return fn;
}

// rank-2 fp32 tensor map: `inner` contiguous elements, `outer` rows `outer_stride` elements apart, 128B swizzle, zero fill
int make_tmap_f32(CUtensorMap* out, const void* base, long long inner, long long outer, long long outer_stride,
                  int box_inner, int box_outer, bool atom32) {
  encode_tiled_fn enc = get_encode_f32();
  DWC_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstr[1] = {(cuuint64_t)outer_stride * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DWC_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(fp32) failed: %d inner=%lld outer=%lld stride=%lld", (int)r, inner,
            outer, outer_stride);
  return 0;
}

template <int BN>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmTcP& p, cudaStream_t st) {
  using Cfg = GtCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    DWC_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  dim3 grid(cdiv(p.M, GT_BM), cdiv(p.N, BN));
  gemm_tf32_kernel<BN><<<grid, GT_THREADS, Cfg::SMEM, st>>>(tmA, tmB, p);
  DWC_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// 1 if (shape, strides, alignment) can run on the tf32 tensor-core kernel
extern "C" int dwc_gemm_tf32_ok(int m, int n, int k, const void* a, int64_t a_sm, int64_t a_sk, const void* b,
                                int64_t b_sk, int64_t b_sn, const void* c, int64_t c_sm, int64_t c_sn) {
  if (m < 64 || n < 32 || k < 32 || c_sn != 1) return 0;
  const bool a_k = a_sk == 1, a_m = a_sm == 1 && !a_k;
  const bool b_k = b_sk == 1, b_n = b_sn == 1 && !b_k;
  if (!(a_k || a_m) || !(b_k || b_n)) return 0;
  const int64_t a_ld = a_k ? a_sm : a_sk, b_ld = b_k ? b_sn : b_sk;
  if ((a_ld % 4) || (b_ld % 4)) return 0;                                      // TMA: strides multiple of 16 bytes
  if (((uintptr_t)a & 15) || ((uintptr_t)b & 15) || ((uintptr_t)c & 3)) return 0;
  return 1;
}

extern "C" int dwc_gemm_tf32(int m, int n, int k, float alpha, const float* a, int64_t a_sm, int64_t a_sk,
                             const float* b, int64_t b_sk, int64_t b_sn, float beta, float* c, int64_t c_sm,
                             const float* bias, int act, dwc_stream_t stream) {
  DWC_CHECK(dwc_gemm_tf32_ok(m, n, k, a, a_sm, a_sk, b, b_sk, b_sn, c, c_sm, 1), "dwc_gemm_tf32: unsupported operands");
  GemmTcP p;
  p.M = m; p.N = n; p.K = k;
  p.a_mn = a_sk == 1 ? 0 : 1;
  p.b_mn = b_sk == 1 ? 0 : 1;
  p.C = c; p.c_sm = c_sm; p.bias = bias; p.alpha = alpha; p.beta = beta; p.act = act;
  // widest tile that still gives the launch a useful number of CTAs
  const int tm = cdiv(m, GT_BM);
  int bn = 256;
  if (n <= 64) bn = 64;
  else if (n <= 128 || tm * cdiv(n, 256) < 48) bn = n <= 128 ? 128 : (tm * cdiv(n, 128) < 48 ? 64 : 128);
  CUtensorMap tmA, tmB;
  if (p.a_mn) {
    if (make_tmap_f32(&tmA, a, m, k, a_sk, 32, 32, true)) return 1;
  } else {
    if (make_tmap_f32(&tmA, a, k, m, a_sm, GT_BK, GT_BM, false)) return 1;
  }
  if (p.b_mn) {
    if (make_tmap_f32(&tmB, b, n, k, b_sk, 32, 32, true)) return 1;
  } else {
    if (make_tmap_f32(&tmB, b, k, n, b_sn, GT_BK, bn, false)) return 1;
  }
  cudaStream_t st = as_stream(stream);
  if (bn == 256) return launch_gemm<256>(tmA, tmB, p, st);
  if (bn == 128) return launch_gemm<128>(tmA, tmB, p, st);
  return launch_gemm<64>(tmA, tmB, p, st);
}
