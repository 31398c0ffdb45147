// Host-side plumbing shared by all translation units: thread-local error text, device
// queries, and TMA tensor-map creation.  cuTensorMapEncodeTiled is looked up through the
// runtime (cudaGetDriverEntryPoint) so the library has no link-time dependency on libcuda
// and can be loaded (symbols checked) on a machine without a GPU driver.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace pevit {

namespace {
thread_local char g_error[1024] = "";
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_error; }

int sm_count() {
  static int cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cached[dev & 63] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev & 63] = n;
  }
  return cached[dev & 63];
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                      uint32_t box_rows, uint32_t box_cols) {
  if (g_encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
      set_error("cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
      return -2;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {row_stride_elems * 2};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estride[2] = {1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu stride=%llu box=%ux%u", (int)r, base,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_stride_elems, box_rows,
              box_cols);
    return -2;
  }
  return 0;
}

}  // namespace pevit
