// Per-warp clock64 event tracer for the attention kernels (diagnostics; tools/attn_trace.py reads the dump).
#pragma once
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace pevit {

// Diagnostics (PEVIT_ATTN_TRACE=<file>): CTA 0 records (clock64 << 8 | event id) per warp into a global buffer that the
// host dumps after the launch; tools/attn_trace.py turns it into a per-role timeline.  Null pointer = off (one
// predictable branch per event).
constexpr int TRACE_EVENTS = 1024;  // per warp
struct Tracer {
  unsigned long long* buf;
  int n;
  __device__ __forceinline__ Tracer(unsigned long long* base, int warp)
      : buf(base != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 ? base + warp * TRACE_EVENTS : nullptr), n(0) {}
  __device__ __forceinline__ void operator()(int id) {
    if (buf != nullptr && n < TRACE_EVENTS) buf[n++] = (static_cast<unsigned long long>(clock64()) << 8) | static_cast<unsigned>(id);
  }
};
struct TraceHost {
  unsigned long long* dev = nullptr;
  const char* path = nullptr;
  int warps = 0;
  unsigned long long* begin(int nwarps) {
    path = getenv("PEVIT_ATTN_TRACE");
    if (path == nullptr) return nullptr;
    warps = nwarps;
    if (cudaMalloc(&dev, sizeof(unsigned long long) * TRACE_EVENTS * nwarps) != cudaSuccess) return dev = nullptr;
    cudaMemset(dev, 0, sizeof(unsigned long long) * TRACE_EVENTS * nwarps);
    return dev;
  }
  void end(cudaStream_t s, const char* tag) {
    if (dev == nullptr) return;
    cudaStreamSynchronize(s);
    const size_t n = static_cast<size_t>(TRACE_EVENTS) * warps;
    unsigned long long* host = static_cast<unsigned long long*>(malloc(n * sizeof(unsigned long long)));
    cudaMemcpy(host, dev, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    char name[512];
    snprintf(name, sizeof(name), "%s.%s.bin", path, tag);
    if (FILE* f = fopen(name, "wb")) { fwrite(&warps, sizeof(int), 1, f); fwrite(host, sizeof(unsigned long long), n, f); fclose(f); }
    free(host);
    cudaFree(dev);
    dev = nullptr;
  }
};

}  // namespace pevit
