// Microbenchmark: tcgen05.ld throughput / latency per SM as a function of active warps and loads in flight.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu ; ./tmem_ld_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ldwait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int INFLIGHT>
__global__ void __launch_bounds__(512, 1) k(int nwarps, int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    uint32_t v[INFLIGHT][32];
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int f = 0; f < INFLIGHT; ++f) ld32(base + ((it * INFLIGHT + f) & 15) * 32, v[f]);
      ldwait();
#pragma unroll
      for (int f = 0; f < INFLIGHT; ++f)
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[f][i];
    }
    t1 = clock64();
  }
  if ((threadIdx.x & 31) == 0 && warp < nwarps) out[warp] = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int F> void run(int nwarps) {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 16 * sizeof(long long)); cudaMalloc(&sink, 512 * 4);
  const int iters = 2000 / F;
  k<F><<<1, 512>>>(nwarps, iters, out, sink);
  long long h[16];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < nwarps; ++i) mx = h[i] > mx ? h[i] : mx;
  const double bytes = double(nwarps) * iters * F * 4096.0;
  printf("warps=%2d inflight=%d: %8lld cycles, %6.1f cyc per load per warp, %7.1f B/clk/SM  (%s)\n", nwarps, F, mx,
         double(mx) / (iters * F), bytes / mx, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(sink);
}
int main() {
  for (int w : {1, 2, 4, 8, 12, 16}) { run<1>(w); run<2>(w); run<3>(w); }
  return 0;
}
