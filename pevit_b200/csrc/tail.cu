// Tail of the fine-tune step: the linear classification head with its cross-entropy loss
// (kadaptation_clip.py:176-185 Classifier.forward -> nn.Linear, :350 CrossEntropyLoss) and the SGD(momentum) update
// (optim/build.py:18-127 -> torch.optim.SGD) over ONE flat parameter / gradient / momentum buffer.  All of it is
// tiny (N x E features, C classes, ~55 k trainable scalars): the point is launch count -- three kernels instead of
// the ~25 element-wise / reduction launches PyTorch issues for the same arithmetic -- and exact fp32 arithmetic.
#include "common.cuh"
#include "kernels.h"

namespace pevit {
namespace {

constexpr int HEAD_THREADS = 128;

// One CTA per sample: logits = feat W^T + b, loss += (logsumexp - logit[label]) / N, dlogits = (softmax - 1hot) / N.
__global__ void __launch_bounds__(HEAD_THREADS)
head_ce_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ W, const float* __restrict__ b,
                   const long long* __restrict__ labels, int N, int E, int C, float* __restrict__ logits,
                   float* __restrict__ dlogits, float* __restrict__ loss) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];  // feat row [E], logits [C]
  float* sf = sm;
  float* sl = sm + E;
  const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < E; e += HEAD_THREADS) sf[e] = feat[static_cast<size_t>(n) * E + e];
  __syncthreads();
  for (int c = warp; c < C; c += HEAD_THREADS / 32) {
    const float* w = W + static_cast<size_t>(c) * E;
    float acc = 0.f;
    for (int e = lane; e < E; e += 32) acc = fmaf(sf[e], w[e], acc);
    acc = warp_sum(acc);
    if (lane == 0) sl[c] = acc + (b != nullptr ? b[c] : 0.f);
  }
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, sl[c]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += __expf(sl[c] - mx);
    se = warp_sum(se);
    const float lse = mx + __logf(se);
    const long long y64 = labels[n];
    // PyTorch's CrossEntropyLoss raises on a class index outside [0, C) (device assert): same contract -- fail the
    // launch loudly instead of dropping the loss term while still emitting a gradient
    if (y64 < 0 || y64 >= C) __trap();
    const int y = static_cast<int>(y64);
    const float inv_n = 1.f / N;
    for (int c = lane; c < C; c += 32) {
      const float p = __expf(sl[c] - lse);
      logits[static_cast<size_t>(n) * C + c] = sl[c];
      dlogits[static_cast<size_t>(n) * C + c] = (p - (c == y ? 1.f : 0.f)) * inv_n;
    }
    if (lane == 0) atomicAdd(loss, (lse - sl[y]) * inv_n);
  }
}

// dfeat[n][e] = g * sum_c dlogits[n][c] W[c][e]  -> bf16 (the A operand of the projection dgrad GEMM)
__global__ void __launch_bounds__(HEAD_THREADS)
head_dfeat_kernel(const float* __restrict__ dlogits, const float* __restrict__ W, const float* __restrict__ gscale, int N,
                  int E, int C, bf16* __restrict__ dfeat) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sd[];  // dlogits row [C]
  const int n = blockIdx.x;
  const float g = gscale != nullptr ? *gscale : 1.f;
  for (int c = threadIdx.x; c < C; c += HEAD_THREADS) sd[c] = dlogits[static_cast<size_t>(n) * C + c] * g;
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += HEAD_THREADS) {
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(sd[c], W[static_cast<size_t>(c) * E + e], acc);
    dfeat[static_cast<size_t>(n) * E + e] = __float2bfloat16(acc);
  }
}

// dW[c][e] (+)= g * sum_n dlogits[n][c] feat[n][e];  db[c] (+)= g * sum_n dlogits[n][c]   (grid: (ceil(E/32), C))
// Block = WG_GROUPS sample groups x 32 feature columns: a thread sums every WG_GROUPS-th sample (a serial loop over all
// N samples per thread was a 256-long chain of dependent L2 loads: 34 us for 1.3 MFLOP), the groups are combined through
// shared memory in a fixed order (deterministic).
constexpr int WG_GROUPS = 8;
__global__ void __launch_bounds__(32 * WG_GROUPS)
head_wgrad_kernel(const float* __restrict__ dlogits, const float* __restrict__ feat, const float* __restrict__ gscale,
                  int N, int E, int C, float* __restrict__ dW, float* __restrict__ db, int accumulate) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float part[WG_GROUPS][33];
  __shared__ float partb[WG_GROUPS];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.y, e = blockIdx.x * 32 + lane;
  const float g = gscale != nullptr ? *gscale : 1.f;
  float acc = 0.f, accb = 0.f;
#pragma unroll 4
  for (int n = grp; n < N; n += WG_GROUPS) {
    const float d = __ldg(dlogits + static_cast<size_t>(n) * C + c);
    if (e < E) acc = fmaf(d, __ldg(feat + static_cast<size_t>(n) * E + e), acc);
    accb += d;
  }
  part[grp][lane] = acc;
  if (lane == 0) partb[grp] = accb;
  __syncthreads();
  if (grp != 0) return;
  float tot = 0.f, totb = 0.f;
#pragma unroll
  for (int q = 0; q < WG_GROUPS; ++q) { tot += part[q][lane]; totb += partb[q]; }
  if (dW != nullptr && e < E) {
    float* dst = dW + static_cast<size_t>(c) * E + e;
    *dst = accumulate ? *dst + g * tot : g * tot;
  }
  if (db != nullptr && blockIdx.x == 0 && lane == 0) db[c] = accumulate ? db[c] + g * totb : g * totb;
}

// torch.optim.SGD (dampening 0, no Nesterov): g' = gscale * g + wd * p;  m = mu * m + g';  p -= lr * m.
// With m zero-initialised the first step gives m = g', which is what torch does when it creates the buffer.
__global__ void sgd_momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, size_t n,
                                    float lr, float mu, float wd, float gscale) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float pi = p[i];
    const float gi = fmaf(wd, pi, gscale * g[i]);
    const float mi = fmaf(mu, m[i], gi);
    m[i] = mi;
    p[i] = fmaf(-lr, mi, pi);
  }
}

}  // namespace

int head_ce_fwd(cudaStream_t s, const float* feat, const float* W, const float* b, const long long* labels, int N, int E,
                int C, float* logits, float* dlogits, float* loss) {
  PEVIT_REQUIRE(N > 0 && E > 0 && C > 0 && (E + C) * sizeof(float) <= 48 * 1024, "head_ce_fwd: bad shape N=%d E=%d C=%d", N, E, C);
  ProfScope prof(s, PC_TAIL);
  PEVIT_CHECK_CUDA(launch_kernel(head_ce_fwd_kernel, dim3(N), dim3(HEAD_THREADS), (E + C) * sizeof(float), s, 1, feat, W, b,
                                 labels, N, E, C, logits, dlogits, loss));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int head_ce_bwd(cudaStream_t s, const float* dlogits, const float* feat, const float* W, const float* gscale, int N, int E,
                int C, bf16* dfeat, float* dW, float* db, int accumulate) {
  PEVIT_REQUIRE(N > 0 && E > 0 && C > 0 && C * sizeof(float) <= 48 * 1024, "head_ce_bwd: bad shape N=%d E=%d C=%d", N, E, C);
  if (dfeat != nullptr) {
    ProfScope prof(s, PC_TAIL);
    PEVIT_CHECK_CUDA(launch_kernel(head_dfeat_kernel, dim3(N), dim3(HEAD_THREADS), C * sizeof(float), s, 1, dlogits, W, gscale,
                                   N, E, C, dfeat));
    PEVIT_CHECK_LAUNCH();
  }
  if (dW != nullptr || db != nullptr) {  // a frozen weight with a trainable bias still needs db
    ProfScope prof(s, PC_TAIL);
    const int gx = dW != nullptr ? (E + 31) / 32 : 1;
    PEVIT_CHECK_CUDA(launch_kernel(head_wgrad_kernel, dim3(gx, C), dim3(32 * WG_GROUPS), 0, s, 1,
                                   dlogits, feat, gscale, N, E, C, dW, db, accumulate));
    PEVIT_CHECK_LAUNCH();
  }
  return 0;
}

int sgd_momentum(cudaStream_t s, float* p, const float* g, float* m, size_t n, float lr, float mu, float wd, float gscale) {
  PEVIT_REQUIRE(p && g && m, "sgd_momentum: null buffer");
  if (n == 0) return 0;
  size_t grid = (n + 255) / 256;
  const size_t cap = static_cast<size_t>(sm_count()) * 4;
  if (grid > cap) grid = cap;
  ProfScope prof(s, PC_TAIL);
  PEVIT_CHECK_CUDA(launch_kernel(sgd_momentum_kernel, dim3(static_cast<unsigned>(grid)), dim3(256), 0, s, 1, p, g, m, n, lr, mu, wd,
                                 gscale));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

}  // namespace pevit
