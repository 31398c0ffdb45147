// Host-side plumbing shared by all translation units: thread-local error text, device
// queries, and TMA tensor-map creation.  cuTensorMapEncodeTiled is looked up through the
// runtime (cudaGetDriverEntryPoint) so the library has no link-time dependency on libcuda
// and can be loaded (symbols checked) on a machine without a GPU driver.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

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

bool pdl_enabled() {
  static const bool on = getenv("PEVIT_PDL") != nullptr;
  return on;
}

// ------------------------------------------------------------------ launch accounting
namespace {
struct ProfEvent { cudaEvent_t start, stop; int cls; };
std::vector<ProfEvent> g_events;   // recorded pairs since the last read
std::vector<ProfEvent> g_pool;     // reusable pairs
std::mutex g_prof_mu;
bool g_prof_on = false;
std::atomic<long long> g_launches{0};
thread_local int g_tag = -1;
const char* const kClassNames[PC_COUNT] = {
    "gemm_qkv", "gemm_out", "gemm_fc", "gemm_proj", "gemm_dproj", "gemm_dfc", "gemm_dout", "gemm_dqkv", "gemm_dT",
    "gemm_delta", "gemm_bottleneck", "gemm_other", "attn_fwd", "attn_bwd", "ln_fwd", "ln_bwd", "atb", "colsum", "expand",
    "factor_grads", "cast", "stem", "gemm_stem", "tail", "allreduce_sgd"};
}  // namespace

void prof_set_tag(int cls) { g_tag = cls; }

ProfScope::ProfScope(cudaStream_t s, int default_cls) : stream(s), slot(-1) {
  const int cls = (g_tag >= 0 && g_tag < PC_COUNT) ? g_tag : default_cls;
  g_tag = -1;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfEvent ev;
  if (!g_pool.empty()) {
    ev = g_pool.back();
    g_pool.pop_back();
  } else if (cudaEventCreate(&ev.start) != cudaSuccess || cudaEventCreate(&ev.stop) != cudaSuccess) {
    return;
  }
  ev.cls = cls;
  cudaEventRecord(ev.start, s);
  g_events.push_back(ev);
  slot = static_cast<int>(g_events.size()) - 1;
}

ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot < static_cast<int>(g_events.size())) cudaEventRecord(g_events[slot].stop, stream);
}

int prof_enable(int on) {
  g_prof_on = on != 0;
  return 0;
}

int prof_reset() {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& e : g_events) g_pool.push_back(e);
  g_events.clear();
  return 0;
}

// Accumulates elapsed ms and launch counts per class (waits for the recorded events), then recycles them.
int prof_read(double* ms, long long* launches, int n) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < n; ++i) { ms[i] = 0.0; launches[i] = 0; }
  for (auto& e : g_events) {
    if (cudaEventSynchronize(e.stop) != cudaSuccess) { set_error("prof_read: event sync failed"); return -2; }
    float t = 0.f;
    if (cudaEventElapsedTime(&t, e.start, e.stop) != cudaSuccess) { set_error("prof_read: elapsed failed"); return -2; }
    if (e.cls < n) { ms[e.cls] += t; launches[e.cls] += 1; }
    g_pool.push_back(e);
  }
  g_events.clear();
  return 0;
}

long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
const char* prof_class_name(int cls) { return (cls >= 0 && cls < PC_COUNT) ? kClassNames[cls] : "?"; }

// cuTensorMapEncodeTiled costs microseconds of host time and the step issues ~200 GEMM launches, most of them on
// the same (pointer, shape) tuples every step (frozen weights always; activations whenever the caching allocator
// hands back the same block).  A descriptor depends only on the tuple hashed here, so cached copies stay valid.
namespace {
struct TmapKey {
  const void* base; uint64_t a, b, c; uint32_t d, e, kind;
  bool operator==(const TmapKey& o) const {
    return base == o.base && a == o.a && b == o.b && c == o.c && d == o.d && e == o.e && kind == o.kind;
  }
};
struct TmapSlot { TmapKey key; CUtensorMap map; bool used; };
constexpr int kTmapSlots = 8192;  // direct-mapped
thread_local TmapSlot* g_tmap_cache = nullptr;
inline size_t tmap_hash(const TmapKey& k) {
  uint64_t h = reinterpret_cast<uint64_t>(k.base) * 0x9E3779B97F4A7C15ull;
  h ^= (k.a + 0x7F4A7C15u) * 0xC2B2AE3D27D4EB4Full; h ^= (k.b << 17) ^ (k.c << 31) ^ (uint64_t(k.d) << 43) ^ (uint64_t(k.e) << 7) ^ k.kind;
  h ^= h >> 29;
  return static_cast<size_t>(h % kTmapSlots);
}
inline bool tmap_lookup(const TmapKey& k, CUtensorMap* out, TmapSlot** slot) {
  if (g_tmap_cache == nullptr) g_tmap_cache = new TmapSlot[kTmapSlots]();
  *slot = &g_tmap_cache[tmap_hash(k)];
  if ((*slot)->used && (*slot)->key == k) { *out = (*slot)->map; return true; }
  return false;
}
inline void tmap_store(TmapSlot* slot, const TmapKey& k, const CUtensorMap& m) { slot->key = k; slot->map = m; slot->used = true; }
}  // namespace

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                      uint32_t box_rows, uint32_t box_cols) {
  const TmapKey key{base, rows, cols, row_stride_elems, box_rows, box_cols, 1};
  TmapSlot* slot = nullptr;
  if (tmap_lookup(key, out, &slot)) return 0;
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
  tmap_store(slot, key, *out);
  return 0;
}

int make_tmap_out_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                     uint32_t box_rows, int elem_bytes) {
  const TmapKey key{base, rows, cols, row_stride_elems, box_rows, static_cast<uint32_t>(elem_bytes), 2};
  TmapSlot* slot = nullptr;
  if (tmap_lookup(key, out, &slot)) return 0;
  CUtensorMap dummy;
  if (g_encode == nullptr && make_tmap_bf16_2d(&dummy, base, 8, 64, 64, 8, 64) != 0) return -2;  // resolves g_encode
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {row_stride_elems * static_cast<cuuint64_t>(elem_bytes)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / elem_bytes), box_rows};
  const cuuint32_t estride[2] = {1, 1};
  CUresult r = g_encode(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(out) failed (%d): base=%p rows=%llu cols=%llu stride=%llu", (int)r, base,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_stride_elems);
    return -2;
  }
  tmap_store(slot, key, *out);
  return 0;
}

int make_tmap_qkv_hm_5d(CUtensorMap* out, const void* base, int L, int NB, int H, uint32_t box_n) {
  const TmapKey key{base, static_cast<uint64_t>(L), static_cast<uint64_t>(NB), static_cast<uint64_t>(H), box_n, 0, 4};
  TmapSlot* slot = nullptr;
  if (tmap_lookup(key, out, &slot)) return 0;
  CUtensorMap dummy;
  if (g_encode == nullptr && make_tmap_bf16_2d(&dummy, base, 8, 64, 64, 8, 64) != 0) return -2;  // resolves g_encode
  const cuuint64_t l = static_cast<cuuint64_t>(L), h = static_cast<cuuint64_t>(H), nb = static_cast<cuuint64_t>(NB);
  const cuuint64_t gdim[5] = {64, l, h, nb, 3};
  const cuuint64_t gstride[4] = {128, l * 128, h * l * 128, nb * h * l * 128};
  const cuuint32_t box[5] = {64, 1, 1, box_n, 1};
  const cuuint32_t estride[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(5d) failed (%d): base=%p L=%d NB=%d H=%d", (int)r, base, L, NB, H);
    return -2;
  }
  tmap_store(slot, key, *out);
  return 0;
}

int make_tmap_bf16_hm3d(CUtensorMap* out, const void* base, int L, int heads, uint32_t box_l) {
  const TmapKey key{base, static_cast<uint64_t>(L), static_cast<uint64_t>(heads), 0, box_l, 0, 5};
  TmapSlot* slot = nullptr;
  if (tmap_lookup(key, out, &slot)) return 0;
  CUtensorMap dummy;
  if (g_encode == nullptr && make_tmap_bf16_2d(&dummy, base, 8, 64, 64, 8, 64) != 0) return -2;  // resolves g_encode
  const cuuint64_t gdim[3] = {64, static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(heads)};
  const cuuint64_t gstride[2] = {128, static_cast<cuuint64_t>(L) * 128};
  const cuuint32_t box[3] = {64, box_l, 1};
  const cuuint32_t estride[3] = {1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed (%d): base=%p L=%d heads=%d box_l=%u", (int)r, base, L, heads, box_l);
    return -2;
  }
  tmap_store(slot, key, *out);
  return 0;
}

int make_tmap_bf16_tok_heads(CUtensorMap* out, const void* base, int L, int NB, int H, uint64_t row_stride_elems,
                             uint32_t box_l) {
  const TmapKey key{base, static_cast<uint64_t>(L), static_cast<uint64_t>(NB), row_stride_elems, static_cast<uint32_t>(H),
                    box_l, 3};
  TmapSlot* slot = nullptr;
  if (tmap_lookup(key, out, &slot)) return 0;
  CUtensorMap dummy;
  if (g_encode == nullptr && make_tmap_bf16_2d(&dummy, base, 8, 64, 64, 8, 64) != 0) return -2;  // resolves g_encode
  const cuuint64_t gdim[4] = {64, static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(NB), static_cast<cuuint64_t>(L)};
  const cuuint64_t gstride[3] = {128, row_stride_elems * 2, static_cast<cuuint64_t>(NB) * row_stride_elems * 2};
  const cuuint32_t box[4] = {64, 1, 1, box_l};
  const cuuint32_t estride[4] = {1, 1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d) failed (%d): base=%p L=%d NB=%d H=%d stride=%llu", (int)r, base, L, NB, H,
              (unsigned long long)row_stride_elems);
    return -2;
  }
  tmap_store(slot, key, *out);
  return 0;
}

}  // namespace pevit
