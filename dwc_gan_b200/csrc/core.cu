// Error plumbing, device queries and TMA descriptor construction (host side).
#include "common.cuh"
#include <stdlib.h>
#include <stdarg.h>
#include <cudaTypedefs.h>

static thread_local char g_err[1024] = "";

void dwc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* dwc_last_error(void) { return g_err; }
extern "C" int dwc_abi_version(void) { return 1; }

int dwc_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DWC_PDL");
    on = e ? atoi(e) : 0;          // measured: +1.2 % step time when on (profiles/r02f), so opt-in
  }
  return on;
}

int dwc_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
  static encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
  }
  return fn;
}

extern "C" int dwc_tc_available(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 && get_encode() != nullptr;
}

int dwc_make_tmap5(CUtensorMap* out, const void* base, const int64_t dim[5], const int64_t str_elems[5], int bx,
                   int by, int bz, int bn) {
  encode_tiled_fn enc = get_encode();
  DWC_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  DWC_CHECK(((uintptr_t)base & 15) == 0, "TMA base address must be 16-byte aligned");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  for (int i = 0; i < 5; ++i) gdim[i] = (cuuint64_t)dim[i];
  for (int i = 1; i < 5; ++i) {
    gstr[i - 1] = (cuuint64_t)str_elems[i] * 2;
    DWC_CHECK(gstr[i - 1] % 16 == 0, "TMA global stride %d (%lld bytes) is not a multiple of 16", i,
              (long long)gstr[i - 1]);
  }
  cuuint32_t box[5] = {64u, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz, (cuuint32_t)bn};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DWC_CHECK(r == CUDA_SUCCESS,
            "cuTensorMapEncodeTiled(rank5) failed: %d dims=(%lld,%lld,%lld,%lld,%lld) box=(64,%d,%d,%d,%d)", (int)r,
            (long long)dim[0], (long long)dim[1], (long long)dim[2], (long long)dim[3], (long long)dim[4], bx, by, bz,
            bn);
  return 0;
}

int dwc_make_tmap2(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t row_stride_elems,
                   int box_rows, int box_cols) {
  encode_tiled_fn enc = get_encode();
  DWC_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  DWC_CHECK(((uintptr_t)base & 15) == 0, "TMA base address must be 16-byte aligned");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)row_stride_elems * 2};
  DWC_CHECK(gstr[0] % 16 == 0, "TMA row stride must be a multiple of 16 bytes");
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DWC_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(rank2) failed: %d rows=%lld cols=%lld box=(%d,%d)", (int)r,
            (long long)rows, (long long)cols, box_rows, box_cols);
  return 0;
}
