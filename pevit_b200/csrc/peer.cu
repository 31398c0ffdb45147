// Data-parallel exchange fused with the optimizer (the reference has no exchange at all, its
// update is torch.optim.SGD built by optim/build.py:18-127 and stepped at kadaptation_clip.py:353).
//
// One launch per step and per rank replaces ncclAllReduce + the SGD kernel(s): a ONE-SHOT all-reduce over peer-mapped
// memory.  Every rank keeps its flat fp32 gradient buffer in a cudaMalloc allocation that all other ranks of the node
// have opened through CUDA IPC, so over NVLink / NVSwitch a CTA simply loads the same 16-byte chunk from all W
// buffers, adds them in rank order (every rank therefore computes bit-identical sums), and applies the momentum-SGD
// update to its own parameters in the same pass.  The buffers are small (KAdaptation ViT-B/32: 55 k floats = 221 KB;
// ViT-L/14: 135 k), so the exchange is latency-bound: one-shot (W reads per element, no reduce-scatter /
// all-gather round trip) is the right shape, and the cost is two flag hand-shakes.
//
// Protocol (per CTA b, epoch e = number of launches so far + 1, kept in device memory so that a captured CUDA graph
// replays it): (1) start barrier -- thread r < W stores e into flag[0][b][me] of rank r (st.release.sys) and spins on
// its own flag[0][b][r] (ld.acquire.sys) until it reads >= e.  A peer's CTA having started means the peer's backward
// kernels (earlier in its stream) are complete, i.e. its gradients are final.  (2) reduce + update.  (3) end barrier
// on flag[1] -- no rank leaves the kernel (and lets the next step's zero-fill touch its gradients) while another rank
// may still be reading them.  Flags are monotonic, so no reset and no ABA.  A spin that exceeds the timeout (20 s;
// PEVIT_PEER_TIMEOUT_MS overrides) gives up,
// raises the error word of the control block and proceeds: a lost peer costs a wrong step, never a hung GPU.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace pevit {
namespace {

constexpr int PEER_THREADS = 512;

struct PeerCtl {
  uint32_t flag[2][PEER_MAX_CTAS][PEER_MAX_WORLD];  // [phase][cta][source rank], written by the source rank
  uint32_t epoch[PEER_MAX_CTAS];                    // local: launches completed by this CTA index
  uint32_t error;                                   // local: 1 = a barrier timed out
};

__host__ __device__ inline size_t ctl_offset(size_t n) { return (n * sizeof(float) + 255) & ~size_t(255); }

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys_f1(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct PeerArgs {
  float* peers[PEER_MAX_WORLD];
  int world, rank;
  size_t n, n_decayed;
  float* p;
  float* m;
  float lr, mu, wd, gscale;
  unsigned long long timeout_ns;
};

// all W ranks: flag[phase][b][me] := e on every rank, then wait until every rank's flag on MY control block is >= e
__device__ __forceinline__ void peer_barrier(const PeerArgs& a, PeerCtl* mine, int phase, int b, uint32_t e) {
  __syncthreads();
  if (threadIdx.x < a.world) {
    const int r = threadIdx.x;
    PeerCtl* theirs = reinterpret_cast<PeerCtl*>(reinterpret_cast<uint8_t*>(a.peers[r]) + ctl_offset(a.n));
    // release at system scope is cumulative: it orders the gradients earlier kernels of this stream wrote (visible to
    // this thread since the kernel boundary) and, through the bar.sync above, this CTA's completed peer loads
    st_release_sys(&theirs->flag[phase][b][a.rank], e);
    const uint32_t* f = &mine->flag[phase][b][r];
    const unsigned long long t0 = global_ns();
    while (static_cast<int32_t>(ld_acquire_sys(f) - e) < 0) {
      if (global_ns() - t0 > a.timeout_ns) {
        mine->error = 1;
        break;
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(PEER_THREADS)
allreduce_sgd_kernel(PeerArgs a) {
  pdl_wait();
  const int b = blockIdx.x, W = a.world;
  PeerCtl* mine = reinterpret_cast<PeerCtl*>(reinterpret_cast<uint8_t*>(a.peers[a.rank]) + ctl_offset(a.n));
  const uint32_t e = mine->epoch[b] + 1;
  peer_barrier(a, mine, 0, b, e);

  // contiguous slice of 4-float chunks per CTA; the sum runs over ranks 0..W-1 in that order on every rank
  const size_t n4 = a.n >> 2;
  const size_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const size_t lo = per * b, hi = lo + per < n4 ? lo + per : n4;
  for (size_t c = lo + threadIdx.x; c < hi; c += PEER_THREADS) {
    // all W peer loads of the chunk are in flight together (one NVLink round trip, not W dependent ones) ...
    float4 v[PEER_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < PEER_MAX_WORLD; ++r)
      if (r < W) v[r] = ld_sys_f4(a.peers[r] + 4 * c);
    // ... and are added in rank order
    float4 g = v[0];
#pragma unroll
    for (int r = 1; r < PEER_MAX_WORLD; ++r)
      if (r < W) { g.x += v[r].x; g.y += v[r].y; g.z += v[r].z; g.w += v[r].w; }
    float4 pv = *reinterpret_cast<float4*>(a.p + 4 * c), mv = *reinterpret_cast<float4*>(a.m + 4 * c);
    const float gs[4] = {g.x, g.y, g.z, g.w};
    float ps[4] = {pv.x, pv.y, pv.z, pv.w}, ms[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float wd = (4 * c + j) < a.n_decayed ? a.wd : 0.f;      // optim/build.py:18-86: decayed group first
      const float gi = fmaf(wd, ps[j], a.gscale * gs[j]);
      ms[j] = fmaf(a.mu, ms[j], gi);
      ps[j] = fmaf(-a.lr, ms[j], ps[j]);
    }
    *reinterpret_cast<float4*>(a.m + 4 * c) = make_float4(ms[0], ms[1], ms[2], ms[3]);
    *reinterpret_cast<float4*>(a.p + 4 * c) = make_float4(ps[0], ps[1], ps[2], ps[3]);
  }
  if (b == 0 && threadIdx.x < (a.n & 3)) {   // tail of n % 4 elements
    const size_t i = (n4 << 2) + threadIdx.x;
    float g = ld_sys_f1(a.peers[0] + i);
    for (int r = 1; r < W; ++r) g += ld_sys_f1(a.peers[r] + i);
    const float wd = i < a.n_decayed ? a.wd : 0.f;
    const float pi = a.p[i];
    const float mi = fmaf(a.mu, a.m[i], fmaf(wd, pi, a.gscale * g));
    a.m[i] = mi;
    a.p[i] = fmaf(-a.lr, mi, pi);
  }

  peer_barrier(a, mine, 1, b, e);
  if (threadIdx.x == 0) mine->epoch[b] = e;
}

}  // namespace

size_t peer_buffer_bytes(size_t n) { return ctl_offset(n) + ((sizeof(PeerCtl) + 255) & ~size_t(255)); }

int peer_alloc(size_t n, void** ptr, void* ipc_handle) {
  PEVIT_REQUIRE(ptr != nullptr && ipc_handle != nullptr && n > 0, "peer_alloc: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == PEER_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  PEVIT_CHECK_CUDA(cudaMalloc(&p, peer_buffer_bytes(n)));
  PEVIT_CHECK_CUDA(cudaMemset(p, 0, peer_buffer_bytes(n)));
  PEVIT_CHECK_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t err = cudaIpcGetMemHandle(&h, p);
  if (err != cudaSuccess) {
    cudaFree(p);
    PEVIT_CHECK_CUDA(err);
  }
  memcpy(ipc_handle, &h, sizeof(h));
  *ptr = p;
  return 0;
}

int peer_open(const void* ipc_handle, void** ptr) {
  PEVIT_REQUIRE(ptr != nullptr && ipc_handle != nullptr, "peer_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  PEVIT_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int peer_close(void* ptr) {
  PEVIT_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int peer_free(void* ptr) {
  PEVIT_CHECK_CUDA(cudaFree(ptr));
  return 0;
}

int peer_status(cudaStream_t s, const void* own, size_t n, int* timed_out) {
  PEVIT_REQUIRE(own != nullptr && timed_out != nullptr, "peer_status: null argument");
  uint32_t err = 0;
  const uint8_t* src = static_cast<const uint8_t*>(own) + ctl_offset(n) + offsetof(PeerCtl, error);
  PEVIT_CHECK_CUDA(cudaMemcpyAsync(&err, src, sizeof(err), cudaMemcpyDeviceToHost, s));
  PEVIT_CHECK_CUDA(cudaStreamSynchronize(s));
  *timed_out = static_cast<int>(err);
  return 0;
}

int allreduce_sgd(cudaStream_t s, void* const* peers, int world, int rank, size_t n, size_t n_decayed, float* p, float* m,
                  float lr, float mu, float wd, float gscale) {
  PEVIT_REQUIRE(peers && p && m, "allreduce_sgd: null buffer");
  PEVIT_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, "allreduce_sgd: rank %d of %d (max %d)",
                rank, world, PEER_MAX_WORLD);
  PEVIT_REQUIRE(n > 0 && n_decayed <= n, "allreduce_sgd: n=%zu n_decayed=%zu", n, n_decayed);
  PEVIT_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (reinterpret_cast<uintptr_t>(m) & 15) == 0,
                "allreduce_sgd: parameter / momentum buffers must be 16-byte aligned");
  PeerArgs a{};
  for (int r = 0; r < world; ++r) {
    PEVIT_REQUIRE(peers[r] != nullptr && (reinterpret_cast<uintptr_t>(peers[r]) & 255) == 0,
                  "allreduce_sgd: peer buffer %d is null or not 256-byte aligned", r);
    a.peers[r] = static_cast<float*>(peers[r]);
  }
  a.world = world; a.rank = rank; a.n = n; a.n_decayed = n_decayed; a.p = p; a.m = m;
  a.lr = lr; a.mu = mu; a.wd = wd; a.gscale = gscale;
  static const long long timeout_ms = getenv("PEVIT_PEER_TIMEOUT_MS") ? atoll(getenv("PEVIT_PEER_TIMEOUT_MS")) : 20000;
  a.timeout_ns = static_cast<unsigned long long>(timeout_ms > 0 ? timeout_ms : 20000) * 1000000ull;
  // latency-bound: a few CTAs (all co-resident with every other rank's CTAs of the same index by construction: the
  // grid is far smaller than the SM count), each owning a contiguous slice
  size_t grid = (n / 4 + PEER_THREADS - 1) / PEER_THREADS;   // about one 16-byte chunk per thread
  if (grid < 1) grid = 1;
  if (grid > PEER_MAX_CTAS) grid = PEER_MAX_CTAS;
  ProfScope prof(s, PC_ALLREDUCE_SGD);
  PEVIT_CHECK_CUDA(launch_kernel(allreduce_sgd_kernel, dim3(static_cast<unsigned>(grid)), dim3(PEER_THREADS), 0, s, 1, a));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

}  // namespace pevit
