// Bottleneck weight preparation for Adapter / Compacter, on the device.
//
// Compacter's PHMLinear (reference evaluation/compacter_model.py:196-308) builds its dense weight every call as
//   H = sum_{i<n} kron(rule_i, left_i right_i),   H[a K + k][c P + p] = sum_i rule[i][a][c] left[i][k] right[i][p]
// (n = 4, rank 1: left_i is (in/n) x 1, right_i is 1 x (out/n)) with a torch.einsum, and autograd walks back through
// it.  Here one launch expands both PHM layers of a block straight into the bf16 operand layouts the bottleneck
// GEMMs read (W = H^T as [out][in] and its transpose), and one launch contracts the dense gradients the block
// backward produces into the factor gradients:
//   dleft[i][k]  = sum_{a,c,p} dH[aK+k][cP+p] rule[i][a][c] right[i][p]
//   dright[i][p] = sum_{a,c,k} dH[aK+k][cP+p] rule[i][a][c] left[i][k]
//   drule[i][a][c] = sum_{k,p} dH[aK+k][cP+p] left[i][k] right[i][p]
// The Adapter's dense down / up weights (adapter_model.py:204-295) only need the bf16 copies + transposes: one launch.
#include "common.cuh"
#include "kernels.h"

namespace pevit {
namespace {

constexpr int PHM_MAXN = 8;

struct PhmLayer {
  const float* left;   // [n][in / n]
  const float* right;  // [n][out / n]
  int in_f, out_f;
  bf16* w;             // [out][in]  = H^T   (B operand of y = x H)
  bf16* w_t;           // [in][out]  = H     (B operand of the dgrad)
};

__global__ void __launch_bounds__(256)
phm_expand_kernel(const float* __restrict__ rule, int n, PhmLayer l0, PhmLayer l1) {
  pdl_launch_dependents();
  pdl_wait();
  const PhmLayer& l = blockIdx.y == 0 ? l0 : l1;
  __shared__ float srule[PHM_MAXN * PHM_MAXN * PHM_MAXN];
  for (int i = threadIdx.x; i < n * n * n; i += blockDim.x) srule[i] = rule[i];
  __syncthreads();
  const int K = l.in_f / n, P = l.out_f / n;
  const int total = l.in_f * l.out_f;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int row = idx / l.out_f, col = idx - row * l.out_f;  // H[row][col]: row over in, col over out
    const int a = row / K, k = row - a * K, c = col / P, p = col - c * P;
    float h = 0.f;
    for (int i = 0; i < n; ++i) h = fmaf(srule[(i * n + a) * n + c] * l.left[i * K + k], l.right[i * P + p], h);
    const bf16 hb = __float2bfloat16(h);
    l.w_t[static_cast<size_t>(row) * l.out_f + col] = hb;
    l.w[static_cast<size_t>(col) * l.in_f + row] = hb;
  }
}

struct PhmGradLayer {
  const float* dh;     // dense gradient, see dh_out_in
  int dh_out_in;       // 0: dh is dH [in][out];  1: dh is dW = dH^T [out][in]
  const float* left;
  const float* right;
  int in_f, out_f;
  float* dleft;
  float* dright;
};

// One CTA per (layer, a, c): the K x P block G_ac = dH[aK.., cP..] of the dense gradient is staged in shared memory once
// (coalesced along whichever index is contiguous in the layer's storage) and contracted there,
//   U_i[k] = sum_p G_ac[k][p] right[i][p],   V_i[p] = sum_k G_ac[k][p] left[i][k],
//   dleft[i][k] += rule[i][a][c] U_i[k],  dright[i][p] += rule[i][a][c] V_i[p],  drule[i][a][c] += sum_k left[i][k] U_i[k];
// the n^2 blocks of a layer (and, for the shared rule, all layers) meet in the outputs through atomicAdd -- the caller
// zeroes them first unless it accumulates.  (The first version ran one warp per output scalar straight from global
// memory: the 128 outputs that contract over the long index walked 96 dependent, uncoalesced loads each -- 33 us per launch.)
constexpr int PFG_THREADS = 256;
__global__ void __launch_bounds__(PFG_THREADS)
phm_factor_grads_kernel(const float* __restrict__ rule, int n, PhmGradLayer l0, PhmGradLayer l1, float* __restrict__ drule) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];
  const PhmGradLayer& l = blockIdx.y == 0 ? l0 : l1;
  const int a = blockIdx.x / n, c = blockIdx.x - a * n;
  const int K = l.in_f / n, P = l.out_f / n, PS = P + 1;
  float* sG = sm;                 // [K][P + 1]
  float* sL = sG + K * PS;        // [n][K]  left factors
  float* sR = sL + n * K;         // [n][P]  right factors
  float* sred = sR + n * P;       // [PFG_THREADS / 32]
  if (l.dh_out_in) {              // stored [out][in]: k is the contiguous index
    for (int e = threadIdx.x; e < K * P; e += PFG_THREADS) {
      const int p = e / K, k = e - p * K;
      sG[k * PS + p] = l.dh[static_cast<size_t>(c * P + p) * l.in_f + a * K + k];
    }
  } else {                        // stored [in][out]: p is the contiguous index
    for (int e = threadIdx.x; e < K * P; e += PFG_THREADS) {
      const int k = e / P, p = e - k * P;
      sG[k * PS + p] = l.dh[static_cast<size_t>(a * K + k) * l.out_f + c * P + p];
    }
  }
  for (int e = threadIdx.x; e < n * K; e += PFG_THREADS) sL[e] = l.left[e];
  for (int e = threadIdx.x; e < n * P; e += PFG_THREADS) sR[e] = l.right[e];
  __syncthreads();
  // U_i has K outputs contracted over P, V_i has P outputs contracted over K; one side of the block is short (<= 32) and
  // one long.  A short contraction is a serial loop of one THREAD per output; a long one takes a WARP per output (lanes
  // stride over the contracted index, shuffle reduction).  (A warp per output for both sides cost 8 k instructions per
  // warp, 39 us; one thread per output for both sides walked 192-long dependent chains.)
  constexpr int NWARPS = PFG_THREADS / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = 0; i < n; ++i) {
    const float rv = rule[(i * n + a) * n + c];
    const float* left = sL + i * K;
    const float* right = sR + i * P;
    float part = 0.f;             // this thread's share of sum_k left[i][k] U_i[k]
    if (P <= 32) {
      for (int k = threadIdx.x; k < K; k += PFG_THREADS) {
        float u = 0.f;
        for (int p = 0; p < P; ++p) u = fmaf(sG[k * PS + p], right[p], u);
        atomicAdd(l.dleft + i * K + k, rv * u);
        part = fmaf(left[k], u, part);
      }
    } else {
      for (int k = warp; k < K; k += NWARPS) {
        float u = 0.f;
        for (int p = lane; p < P; p += 32) u = fmaf(sG[k * PS + p], right[p], u);
        u = warp_sum(u);
        if (lane == 0) {
          atomicAdd(l.dleft + i * K + k, rv * u);
          part = fmaf(left[k], u, part);
        }
      }
    }
    if (K <= 32) {
      for (int p = threadIdx.x; p < P; p += PFG_THREADS) {
        float v = 0.f;
        for (int k = 0; k < K; ++k) v = fmaf(sG[k * PS + p], left[k], v);
        atomicAdd(l.dright + i * P + p, rv * v);
      }
    } else {
      for (int p = warp; p < P; p += NWARPS) {
        float v = 0.f;
        for (int k = lane; k < K; k += 32) v = fmaf(sG[k * PS + p], left[k], v);
        v = warp_sum(v);
        if (lane == 0) atomicAdd(l.dright + i * P + p, rv * v);
      }
    }
    if (drule != nullptr) {
      part = warp_sum(part);
      if (lane == 0) sred[warp] = part;
      __syncthreads();
      if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < NWARPS; ++w) t += sred[w];
        atomicAdd(drule + (i * n + a) * n + c, t);   // both layers (and every block of the model) add into the shared rule
      }
      __syncthreads();
    }
  }
}

// Adapter: dense fp32 down [B][D] / up [D][B] -> bf16 operands and their transposes in one launch
__global__ void __launch_bounds__(256)
bottleneck_pack_kernel(const float* __restrict__ w_down, const float* __restrict__ w_up, int D, int B, bf16* __restrict__ o_down,
                       bf16* __restrict__ o_down_t, bf16* __restrict__ o_up, bf16* __restrict__ o_up_t) {
  pdl_launch_dependents();
  pdl_wait();
  const int total = D * B;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * total; idx += gridDim.x * blockDim.x) {
    if (idx < total) {  // w_down[b][d]
      const int b = idx / D, d = idx - b * D;
      const bf16 v = __float2bfloat16(w_down[idx]);
      o_down[idx] = v;
      o_down_t[static_cast<size_t>(d) * B + b] = v;
    } else {            // w_up[d][b]
      const int j = idx - total;
      const int d = j / B, b = j - d * B;
      const bf16 v = __float2bfloat16(w_up[j]);
      o_up[j] = v;
      o_up_t[static_cast<size_t>(b) * D + d] = v;
    }
  }
}

}  // namespace

int phm_expand(cudaStream_t s, const float* rule, int n, const float* down_left, const float* down_right,
               const float* up_left, const float* up_right, int D, int B, bf16* w_down, bf16* w_down_t, bf16* w_up,
               bf16* w_up_t) {
  PEVIT_REQUIRE(n >= 1 && n <= PHM_MAXN && D % n == 0 && B % n == 0, "phm_expand: n=%d must divide D=%d and %d", n, D, B);
  PhmLayer dn{down_left, down_right, D, B, w_down, w_down_t};
  PhmLayer up{up_left, up_right, B, D, w_up, w_up_t};
  const int grid = (D * B + 255) / 256;
  ProfScope prof(s, PC_EXPAND);
  PEVIT_CHECK_CUDA(launch_kernel(phm_expand_kernel, dim3(grid, 2), dim3(256), 0, s, 1, rule, n, dn, up));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int phm_factor_grads(cudaStream_t s, const float* d_w_down, const float* d_w_up, const float* rule, int n,
                     const float* down_left, const float* down_right, const float* up_left, const float* up_right, int D,
                     int B, float* d_rule, float* d_down_left, float* d_down_right, float* d_up_left, float* d_up_right,
                     bool accumulate) {
  PEVIT_REQUIRE(n >= 1 && n <= PHM_MAXN && D % n == 0 && B % n == 0, "phm_factor_grads: n=%d must divide D=%d and %d", n, D, B);
  // block backward layouts: d_w_down is [D][B] = dH_down ([in][out]); d_w_up is [D][B] = dW_up = dH_up^T ([out][in])
  PhmGradLayer dn{d_w_down, 0, down_left, down_right, D, B, d_down_left, d_down_right};
  PhmGradLayer up{d_w_up, 1, up_left, up_right, B, D, d_up_left, d_up_right};
  if (!accumulate) {  // the kernel adds (n^2 CTAs per layer meet in every output): plain assignment = zero first
    PEVIT_CHECK_CUDA(cudaMemsetAsync(d_down_left, 0, sizeof(float) * D, s));
    PEVIT_CHECK_CUDA(cudaMemsetAsync(d_down_right, 0, sizeof(float) * B, s));
    PEVIT_CHECK_CUDA(cudaMemsetAsync(d_up_left, 0, sizeof(float) * B, s));
    PEVIT_CHECK_CUDA(cudaMemsetAsync(d_up_right, 0, sizeof(float) * D, s));
  }
  const int K = D / n, P = B / n;   // the up layer has the same block size, transposed
  const size_t smem = (static_cast<size_t>(K > P ? K : P) * ((K > P ? P : K) + 1) + static_cast<size_t>(n) * (K + P) + PFG_THREADS / 32) * sizeof(float);
  PEVIT_REQUIRE(smem <= 48 * 1024, "phm_factor_grads: block %d x %d does not fit in shared memory", K, P);
  ProfScope prof(s, PC_FACTOR_GRADS);
  PEVIT_CHECK_CUDA(launch_kernel(phm_factor_grads_kernel, dim3(n * n, 2), dim3(PFG_THREADS), smem, s, 1, rule, n, dn, up, d_rule));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int bottleneck_pack(cudaStream_t s, const float* w_down, const float* w_up, int D, int B, bf16* o_down, bf16* o_down_t,
                    bf16* o_up, bf16* o_up_t) {
  ProfScope prof(s, PC_EXPAND);
  PEVIT_CHECK_CUDA(launch_kernel(bottleneck_pack_kernel, dim3((2 * D * B + 255) / 256), dim3(256), 0, s, 1, w_down, w_up, D, B,
                                 o_down, o_down_t, o_up, o_up_t));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

}  // namespace pevit
