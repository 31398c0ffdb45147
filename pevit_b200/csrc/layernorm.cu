// LayerNorm forward / backward for the fp32 residual stream (reference evaluation/model.py:154-160:
// statistics in fp32, eps 1e-5).  One warp per token row; the row lives in registers between the
// statistics pass and the normalise pass, so x is read from HBM exactly once.  Forward emits the
// bf16 copy that feeds the next tcgen05 GEMM's TMA loads; backward emits fp32 dx (+ residual
// gradient) and the bf16 copy that is the A operand of the next dgrad GEMM.
#include "common.cuh"
#include "kernels.h"

namespace pevit {
namespace {

constexpr int LN_THREADS = 256;
constexpr int LN_ROWS = LN_THREADS / 32;
constexpr int MAXV = 8;  // float4 per lane: D <= 1024 (ViT-B 768, ViT-L 1024)

// NV = D / 128 float4 per lane (compile-time for the ViT widths so the row lives in exactly NV registers
// quads).  Persistent: a fixed grid of warps strides over the rows, so there is no per-row block launch
// and the loads of many rows are in flight per SM.
template <int NV, bool kF32Out, bool kBf16Out>
__global__ void __launch_bounds__(LN_THREADS)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              bf16* __restrict__ y_bf16, float* __restrict__ y_f32, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, int M, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_d = 1.f / D;
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  for (int row = blockIdx.x * LN_ROWS + warp; row < M; row += gridDim.x * LN_ROWS) {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
    float4 buf[NV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) buf[i] = xr[lane + 32 * i];
#pragma unroll
    for (int i = 0; i < NV; ++i) sum += (buf[i].x + buf[i].y) + (buf[i].z + buf[i].w);
    const float mean = warp_sum(sum) * inv_d;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float a = buf[i].x - mean, b = buf[i].y - mean, c = buf[i].z - mean, d = buf[i].w - mean;
      var += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(var) * inv_d + 1e-5f);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
      float4 o;
      o.x = (buf[i].x - mean) * rstd * g.x + b.x;
      o.y = (buf[i].y - mean) * rstd * g.y + b.y;
      o.z = (buf[i].z - mean) * rstd * g.z + b.z;
      o.w = (buf[i].w - mean) * rstd * g.w + b.w;
      if (kF32Out) reinterpret_cast<float4*>(y_f32 + static_cast<size_t>(row) * D)[c] = o;
      if (kBf16Out)
        reinterpret_cast<uint2*>(y_bf16 + static_cast<size_t>(row) * D)[c] =
            make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) [+ dres],  g = dyn * gamma, xhat = (x-mean)*rstd
// kBf16In: dyn is the bf16 output of the dgrad GEMM that produced it (half the traffic of an fp32 hand-over)
template <int NV, bool kParamGrads, bool kBf16In>
__global__ void __launch_bounds__(LN_THREADS)
ln_bwd_kernel(const float* __restrict__ dyn, const float* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ dres,
              float* __restrict__ dx, bf16* __restrict__ dx_bf16, float* __restrict__ dgamma,
              float* __restrict__ dbeta, int M, int D, int dres_rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_d = 1.f / D;
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float4 accg[kParamGrads ? NV : 1], accb[kParamGrads ? NV : 1];
  if (kParamGrads) {
#pragma unroll
    for (int i = 0; i < NV; ++i) accg[i] = accb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int row = blockIdx.x * LN_ROWS + warp; row < M; row += gridDim.x * LN_ROWS) {
    const size_t base = static_cast<size_t>(row) * D;
    const float4* dr = reinterpret_cast<const float4*>(dyn + base);
    const uint2* dr16 = reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(dyn) + base);
    const float4* xr = reinterpret_cast<const float4*>(x + base);
    const float4* rr = reinterpret_cast<const float4*>(dres + base);
    float4 g[NV], xh[NV], res[NV];
    // all loads of the row are issued before the first use
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if constexpr (kBf16In) {
        const uint2 u = dr16[lane + 32 * i];
        const float2 lo = unpack_bf16(u.x), hi = unpack_bf16(u.y);
        g[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
      } else {
        g[i] = dr[lane + 32 * i];
      }
      xh[i] = xr[lane + 32 * i];
    }
    const bool has_res = dres != nullptr && row < dres_rows;
    if (has_res) {
#pragma unroll
      for (int i = 0; i < NV; ++i) res[i] = rr[lane + 32 * i];
    }
    const float mean = mean_in[row], rstd = rstd_in[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 gm = __ldg(g4 + lane + 32 * i);
      const float4 d = g[i];
      xh[i] = make_float4((xh[i].x - mean) * rstd, (xh[i].y - mean) * rstd, (xh[i].z - mean) * rstd,
                          (xh[i].w - mean) * rstd);
      g[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
      if (kParamGrads) {
        accg[i].x += d.x * xh[i].x; accg[i].y += d.y * xh[i].y; accg[i].z += d.z * xh[i].z; accg[i].w += d.w * xh[i].w;
        accb[i].x += d.x; accb[i].y += d.y; accb[i].z += d.z; accb[i].w += d.w;
      }
    }
    const float m1 = warp_sum(s1) * inv_d, m2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      float4 o;
      o.x = rstd * (g[i].x - m1 - xh[i].x * m2);
      o.y = rstd * (g[i].y - m1 - xh[i].y * m2);
      o.z = rstd * (g[i].z - m1 - xh[i].z * m2);
      o.w = rstd * (g[i].w - m1 - xh[i].w * m2);
      if (has_res) { o.x += res[i].x; o.y += res[i].y; o.z += res[i].z; o.w += res[i].w; }
      if (dx != nullptr) reinterpret_cast<float4*>(dx + base)[c] = o;
      if (dx_bf16 != nullptr)
        reinterpret_cast<uint2*>(dx_bf16 + base)[c] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
  }
  if (kParamGrads) {
    // every lane owns fixed columns: the CTA's warps first combine in shared memory, then ONE global atomic per
    // (CTA, column) -- per-warp global atomics put thousands of contenders on each of the 2*D addresses
    __shared__ float sacc[2 * 128 * MAXV];
    for (int c = threadIdx.x; c < 2 * D; c += LN_THREADS) sacc[c] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 4;
      atomicAdd(&sacc[c], accg[i].x); atomicAdd(&sacc[c + 1], accg[i].y);
      atomicAdd(&sacc[c + 2], accg[i].z); atomicAdd(&sacc[c + 3], accg[i].w);
      atomicAdd(&sacc[D + c], accb[i].x); atomicAdd(&sacc[D + c + 1], accb[i].y);
      atomicAdd(&sacc[D + c + 2], accb[i].z); atomicAdd(&sacc[D + c + 3], accb[i].w);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += LN_THREADS) {
      atomicAdd(dgamma + c, sacc[c]);
      atomicAdd(dbeta + c, sacc[D + c]);
    }
  }
}

int ln_grid(int M, int blocks_per_sm) {
  const int need = (M + LN_ROWS - 1) / LN_ROWS;
  const int cap = sm_count() * blocks_per_sm;
  return need < cap ? need : cap;
}

template <int NV>
int launch_fwd(cudaStream_t s, const float* x, const float* gamma, const float* beta, bf16* y_bf16, float* y_f32,
               float* mean, float* rstd, int M, int D) {
  const int grid = ln_grid(M, 6);
  if (y_bf16 && y_f32)
    launch_kernel(ln_fwd_kernel<NV, true, true>, dim3(grid), dim3(LN_THREADS), 0, s, 1, x, gamma, beta, y_bf16, y_f32, mean, rstd, M, D);
  else if (y_bf16)
    launch_kernel(ln_fwd_kernel<NV, false, true>, dim3(grid), dim3(LN_THREADS), 0, s, 1, x, gamma, beta, y_bf16, y_f32, mean, rstd, M, D);
  else
    launch_kernel(ln_fwd_kernel<NV, true, false>, dim3(grid), dim3(LN_THREADS), 0, s, 1, x, gamma, beta, y_bf16, y_f32, mean, rstd, M, D);
  return 0;
}

template <int NV>
int launch_bwd(cudaStream_t s, const float* dyn, const float* x, const float* gamma, const float* mean,
               const float* rstd, const float* dres, float* dx, bf16* dx_bf16, float* dgamma, float* dbeta, int M,
               int D, int dres_rows, bool bf16_in) {
  if (bf16_in)
    launch_kernel(ln_bwd_kernel<NV, false, true>, dim3(ln_grid(M, 4)), dim3(LN_THREADS), 0, s, 1, dyn, x, gamma, mean, rstd, dres,
                  dx, dx_bf16, static_cast<float*>(nullptr), static_cast<float*>(nullptr), M, D, dres_rows);
  else if (dgamma != nullptr)
    launch_kernel(ln_bwd_kernel<NV, true, false>, dim3(ln_grid(M, 2)), dim3(LN_THREADS), 0, s, 1, dyn, x, gamma, mean, rstd, dres, dx,
                  dx_bf16, dgamma, dbeta, M, D, dres_rows);
  else
    launch_kernel(ln_bwd_kernel<NV, false, false>, dim3(ln_grid(M, 4)), dim3(LN_THREADS), 0, s, 1, dyn, x, gamma, mean, rstd, dres, dx,
                  dx_bf16, static_cast<float*>(nullptr), static_cast<float*>(nullptr), M, D, dres_rows);
  return 0;
}

}  // namespace

int layernorm_fwd(cudaStream_t s, const float* x, const float* gamma, const float* beta, bf16* y_bf16, float* y_f32,
                  float* mean, float* rstd, int M, int D) {
  PEVIT_REQUIRE(D % 128 == 0 && D <= 128 * MAXV, "layernorm: D=%d must be a multiple of 128 and <= %d", D, 128 * MAXV);
  PEVIT_REQUIRE(y_bf16 != nullptr || y_f32 != nullptr, "layernorm_fwd: no output");
  ProfScope prof(s, PC_LN_FWD);
  switch (D / 128) {
    case 1: launch_fwd<1>(s, x, gamma, beta, y_bf16, y_f32, mean, rstd, M, D); break;
    case 2: launch_fwd<2>(s, x, gamma, beta, y_bf16, y_f32, mean, rstd, M, D); break;
    case 3: launch_fwd<3>(s, x, gamma, beta, y_bf16, y_f32, mean, rstd, M, D); break;
    case 4: launch_fwd<4>(s, x, gamma, beta, y_bf16, y_f32, mean, rstd, M, D); break;
    case 5: launch_fwd<5>(s, x, gamma, beta, y_bf16, y_f32, mean, rstd, M, D); break;
    case 6: launch_fwd<6>(s, x, gamma, beta, y_bf16, y_f32, mean, rstd, M, D); break;
    case 7: launch_fwd<7>(s, x, gamma, beta, y_bf16, y_f32, mean, rstd, M, D); break;
    default: launch_fwd<8>(s, x, gamma, beta, y_bf16, y_f32, mean, rstd, M, D); break;
  }
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int layernorm_bwd(cudaStream_t s, const float* dyn, const float* x, const float* gamma, const float* mean,
                  const float* rstd, const float* dres, float* dx, bf16* dx_bf16, float* dgamma, float* dbeta, int M,
                  int D, int dres_rows, const bf16* dyn_bf16) {
  PEVIT_REQUIRE(D % 128 == 0 && D <= 128 * MAXV, "layernorm: D=%d must be a multiple of 128 and <= %d", D, 128 * MAXV);
  PEVIT_REQUIRE(dgamma == nullptr || dbeta != nullptr, "layernorm_bwd: dgamma without dbeta");
  PEVIT_REQUIRE(dyn_bf16 == nullptr || dgamma == nullptr, "layernorm_bwd: bf16 dyn is for frozen-gamma LayerNorms only");
  const bool bf16_in = dyn_bf16 != nullptr;
  if (bf16_in) dyn = reinterpret_cast<const float*>(dyn_bf16);
  if (dres_rows < 0) dres_rows = M;
  ProfScope prof(s, PC_LN_BWD);
  switch (D / 128) {
    case 1: launch_bwd<1>(s, dyn, x, gamma, mean, rstd, dres, dx, dx_bf16, dgamma, dbeta, M, D, dres_rows, bf16_in); break;
    case 2: launch_bwd<2>(s, dyn, x, gamma, mean, rstd, dres, dx, dx_bf16, dgamma, dbeta, M, D, dres_rows, bf16_in); break;
    case 3: launch_bwd<3>(s, dyn, x, gamma, mean, rstd, dres, dx, dx_bf16, dgamma, dbeta, M, D, dres_rows, bf16_in); break;
    case 4: launch_bwd<4>(s, dyn, x, gamma, mean, rstd, dres, dx, dx_bf16, dgamma, dbeta, M, D, dres_rows, bf16_in); break;
    case 5: launch_bwd<5>(s, dyn, x, gamma, mean, rstd, dres, dx, dx_bf16, dgamma, dbeta, M, D, dres_rows, bf16_in); break;
    case 6: launch_bwd<6>(s, dyn, x, gamma, mean, rstd, dres, dx, dx_bf16, dgamma, dbeta, M, D, dres_rows, bf16_in); break;
    case 7: launch_bwd<7>(s, dyn, x, gamma, mean, rstd, dres, dx, dx_bf16, dgamma, dbeta, M, D, dres_rows, bf16_in); break;
    default: launch_bwd<8>(s, dyn, x, gamma, mean, rstd, dres, dx, dx_bf16, dgamma, dbeta, M, D, dres_rows, bf16_in); break;
  }
  PEVIT_CHECK_LAUNCH();
  return 0;
}

}  // namespace pevit
