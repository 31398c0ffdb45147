// Attention core with the in-kernel low-rank delta (KAdaptation / LoRA), CUDA-core version.
// One CTA per (image, head).  This is the straightforward on-device statement of
// reference evaluation/model.py:786-815 (and lora_model.py:719-733): q' = q/8 + scr(dq),
// v' = v + scr(dv) with scr = the raw reshape of F4, S = q'k^T, P = softmax(S), O = P v'.
// The delta is expanded in-kernel from the rank-r activations T and the small factor matrix,
// so neither H (DxD) nor the (L*N x D) delta ever exists in HBM.
//
// It is the first correct CUDA path and stays as the on-device cross-check for the
// tcgen05 attention pipeline (attention_tc.cu).
#include "common.cuh"
#include "kernels.h"

namespace pevit {
namespace {

constexpr int AT_THREADS = 128;
constexpr int AT_WARPS = AT_THREADS / 32;
constexpr int QS = 65;   // fp32 row stride (conflict-free column walks)
constexpr int KS = 66;   // bf16 row stride (33 words)
constexpr int MAXT = 9;  // keys per lane: L <= 288

// delta value for element (l, d) of head h of image n (F4 scramble):
// flat chunk j = h*L + l of the image's (L x D) delta -> row n*L + j/H, column block j%H.
__device__ __forceinline__ void delta_rowcol(int n, int h, int l, int L, int H, int& row, int& colblk) {
  const int j = h * L + l;
  row = n * L + j / H;
  colblk = j % H;
}

__device__ __forceinline__ float delta_elem(const bf16* __restrict__ Trow, const float* __restrict__ Qrow, int r,
                                            float alpha) {
  float acc = 0.f;
  for (int i = 0; i < r; ++i) acc = fmaf(__bfloat162float(Trow[i]), __ldg(Qrow + i), acc);
  return alpha * acc;
}

// Loads q', k, v' of one (n, h) into shared memory.  sQ/sV fp32 [L][QS]; sK bf16 [L][KS].
__device__ void load_head(const AttnShape& a, int n, int h, const bf16* __restrict__ q, const bf16* __restrict__ k,
                          const bf16* __restrict__ v, const bf16* __restrict__ T, const float* __restrict__ Qmat,
                          const float* __restrict__ bias, float* sQ, bf16* sK, float* sV) {
  const int L = a.L, H = a.H, D = a.D, r = a.r;
  const size_t head_off = (static_cast<size_t>(n) * H + h) * L * 64;
  for (int idx = threadIdx.x; idx < L * 64; idx += AT_THREADS) {
    const int l = idx >> 6, d = idx & 63;
    float qv = __bfloat162float(q[head_off + idx]);
    float vv = __bfloat162float(v[head_off + idx]);
    sK[l * KS + d] = k[head_off + idx];
    if (r > 0 || bias != nullptr) {
      int row, cb;
      delta_rowcol(n, h, l, L, H, row, cb);
      const int col = cb * 64 + d;
      float dq = 0.f, dv = 0.f;
      if (r > 0) {
        const bf16* Trow = T + static_cast<size_t>(row) * 2 * r;
        dq = delta_elem(Trow, Qmat + static_cast<size_t>(col) * r, r, a.alpha);
        dv = delta_elem(Trow + r, Qmat + (static_cast<size_t>(D) + col) * r, r, a.alpha);
      }
      if (bias != nullptr) { const float b = __ldg(bias + col); dq += b; dv += b; }
      qv += dq;
      vv += dv;
    }
    sQ[l * QS + d] = qv;
    sV[l * QS + d] = vv;
  }
}

__global__ void __launch_bounds__(AT_THREADS)
attn_fwd_ref_kernel(AttnShape a, const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                    const bf16* __restrict__ T, const float* __restrict__ Qmat, const float* __restrict__ bias,
                    bf16* __restrict__ o_tok, float* __restrict__ lse) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int L = a.L, H = a.H;
  float* sQ = reinterpret_cast<float*>(smem);
  float* sV = sQ + L * QS;
  float* sP = sV + L * QS;                                // [AT_WARPS][L]
  bf16* sK = reinterpret_cast<bf16*>(sP + AT_WARPS * L);  // [L][KS]
  const int n = blockIdx.x / H, h = blockIdx.x % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_head(a, n, h, q, k, v, T, Qmat, bias, sQ, sK, sV);
  __syncthreads();
  float* myP = sP + warp * L;
  for (int i = warp; i < L; i += AT_WARPS) {
    float s[MAXT];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int j = lane + 32 * t;
      s[t] = -INFINITY;
      if (j < L) {
        float acc = 0.f;
        const __nv_bfloat162* kr = reinterpret_cast<const __nv_bfloat162*>(sK + j * KS);
#pragma unroll 8
        for (int d = 0; d < 32; ++d) {
          const float2 kk = __bfloat1622float2(kr[d]);
          acc = fmaf(sQ[i * QS + 2 * d], kk.x, acc);
          acc = fmaf(sQ[i * QS + 2 * d + 1], kk.y, acc);
        }
        s[t] = acc;
        mx = fmaxf(mx, acc);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int j = lane + 32 * t;
      if (j < L) { s[t] = __expf(s[t] - mx); sum += s[t]; }
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int j = lane + 32 * t;
      if (j < L) myP[j] = s[t] * inv;
    }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < L; ++j) {
      const float p = myP[j];
      o0 = fmaf(p, sV[j * QS + lane], o0);
      o1 = fmaf(p, sV[j * QS + lane + 32], o1);
    }
    bf16* orow = o_tok + (static_cast<size_t>(i) * a.NB + n) * a.D + h * 64;
    orow[lane] = __float2bfloat16(o0);
    orow[lane + 32] = __float2bfloat16(o1);
    if (lane == 0) lse[(static_cast<size_t>(n) * H + h) * L + i] = mx + __logf(sum);
    __syncwarp();
  }
}

__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_ref_kernel(AttnShape a, const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                    const bf16* __restrict__ T, const float* __restrict__ Qmat, const float* __restrict__ bias,
                    const bf16* __restrict__ o_tok, const bf16* __restrict__ do_tok, const float* __restrict__ lse,
                    bf16* __restrict__ dqkv, int ld, bf16* __restrict__ ddelta) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int L = a.L, H = a.H, D = a.D, NB = a.NB;
  float* sQ = reinterpret_cast<float*>(smem);
  float* sV = sQ + L * QS;
  float* sA = sV + L * QS;          // [AT_WARPS][L]  p   (or p_ij over i)
  float* sB = sA + AT_WARPS * L;    // [AT_WARPS][L]  ds
  float* sLse = sB + AT_WARPS * L;  // [L]
  float* sDel = sLse + L;           // [L] rowsum(dO * O)
  bf16* sK = reinterpret_cast<bf16*>(sDel + L);  // [L][KS]
  bf16* sdO = sK + L * KS;                       // [L][KS]
  const int n = blockIdx.x / H, h = blockIdx.x % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t head_off = (static_cast<size_t>(n) * H + h) * L * 64;

  load_head(a, n, h, q, k, v, T, Qmat, bias, sQ, sK, sV);
  for (int idx = threadIdx.x; idx < L * 64; idx += AT_THREADS) {
    const int l = idx >> 6, d = idx & 63;
    sdO[l * KS + d] = do_tok[(static_cast<size_t>(l) * NB + n) * D + h * 64 + d];
  }
  for (int i = threadIdx.x; i < L; i += AT_THREADS) sLse[i] = lse[(static_cast<size_t>(n) * H + h) * L + i];
  __syncthreads();
  float* myA = sA + warp * L;
  float* myB = sB + warp * L;
  // ---- pass A: one warp per query row i -> dQ'[i].  delta_i = sum_j P_ij dP_ij is taken from the
  // very P and dP that form dS (identical to rowsum(dO * O) in exact arithmetic, but free of the
  // bf16 rounding of O, so sum_j dS_ij stays ~0) and is kept in shared memory for pass B.
  for (int i = warp; i < L; i += AT_WARPS) {
    const float lse_i = sLse[i];
    float pv[MAXT], dpv[MAXT];
    float del = 0.f;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int j = lane + 32 * t;
      pv[t] = 0.f; dpv[t] = 0.f;
      if (j < L) {
        float sc = 0.f, dp = 0.f;
        const __nv_bfloat162* kr = reinterpret_cast<const __nv_bfloat162*>(sK + j * KS);
        const __nv_bfloat162* dor = reinterpret_cast<const __nv_bfloat162*>(sdO + i * KS);
#pragma unroll 8
        for (int d = 0; d < 32; ++d) {
          const float2 kk = __bfloat1622float2(kr[d]);
          const float2 dd = __bfloat1622float2(dor[d]);
          sc = fmaf(sQ[i * QS + 2 * d], kk.x, sc);
          sc = fmaf(sQ[i * QS + 2 * d + 1], kk.y, sc);
          dp = fmaf(dd.x, sV[j * QS + 2 * d], dp);
          dp = fmaf(dd.y, sV[j * QS + 2 * d + 1], dp);
        }
        pv[t] = __expf(sc - lse_i);
        dpv[t] = dp;
        del = fmaf(pv[t], dp, del);
      }
    }
    del = warp_sum(del);
    if (lane == 0) sDel[i] = del;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int j = lane + 32 * t;
      if (j < L) myB[j] = pv[t] * (dpv[t] - del);
    }
    __syncwarp();
    float g0 = 0.f, g1 = 0.f;
    for (int j = 0; j < L; ++j) {
      const float ds = myB[j];
      g0 = fmaf(ds, __bfloat162float(sK[j * KS + lane]), g0);
      g1 = fmaf(ds, __bfloat162float(sK[j * KS + lane + 32]), g1);
    }
    bf16* dq_tok = dqkv + (static_cast<size_t>(i) * NB + n) * ld + h * 64;
    dq_tok[lane] = __float2bfloat16(g0 * 0.125f);
    dq_tok[lane + 32] = __float2bfloat16(g1 * 0.125f);
    if (ddelta != nullptr) {
      ddelta[head_off + i * 64 + lane] = __float2bfloat16(g0);
      ddelta[head_off + i * 64 + lane + 32] = __float2bfloat16(g1);
    }
    __syncwarp();
  }
  __syncthreads();  // sDel complete
  // ---- pass B: one warp per key row j -> dK[j], dV'[j]
  const size_t dv_plane = static_cast<size_t>(NB) * H * L * 64;
  for (int j = warp; j < L; j += AT_WARPS) {
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int i = lane + 32 * t;
      if (i < L) {
        float sc = 0.f, dp = 0.f;
        const __nv_bfloat162* kr = reinterpret_cast<const __nv_bfloat162*>(sK + j * KS);
        const __nv_bfloat162* dor = reinterpret_cast<const __nv_bfloat162*>(sdO + i * KS);
#pragma unroll 8
        for (int d = 0; d < 32; ++d) {
          const float2 kk = __bfloat1622float2(kr[d]);
          const float2 dd = __bfloat1622float2(dor[d]);
          sc = fmaf(sQ[i * QS + 2 * d], kk.x, sc);
          sc = fmaf(sQ[i * QS + 2 * d + 1], kk.y, sc);
          dp = fmaf(dd.x, sV[j * QS + 2 * d], dp);
          dp = fmaf(dd.y, sV[j * QS + 2 * d + 1], dp);
        }
        const float p = __expf(sc - sLse[i]);
        myA[i] = p;
        myB[i] = p * (dp - sDel[i]);
      }
    }
    __syncwarp();
    float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int i = 0; i < L; ++i) {
      const float p = myA[i], ds = myB[i];
      k0 = fmaf(ds, sQ[i * QS + lane], k0);
      k1 = fmaf(ds, sQ[i * QS + lane + 32], k1);
      v0 = fmaf(p, __bfloat162float(sdO[i * KS + lane]), v0);
      v1 = fmaf(p, __bfloat162float(sdO[i * KS + lane + 32]), v1);
    }
    bf16* tok = dqkv + (static_cast<size_t>(j) * NB + n) * ld + h * 64;
    tok[D + lane] = __float2bfloat16(k0);
    tok[D + lane + 32] = __float2bfloat16(k1);
    tok[2 * D + lane] = __float2bfloat16(v0);
    tok[2 * D + lane + 32] = __float2bfloat16(v1);
    if (ddelta != nullptr) {
      ddelta[dv_plane + head_off + j * 64 + lane] = __float2bfloat16(v0);
      ddelta[dv_plane + head_off + j * 64 + lane + 32] = __float2bfloat16(v1);
    }
    __syncwarp();
  }
}

int check_shape(const AttnShape& a) {
  PEVIT_REQUIRE(a.L > 0 && a.L <= 32 * MAXT, "attention: L=%d out of range (1..%d)", a.L, 32 * MAXT);
  PEVIT_REQUIRE(a.H * 64 == a.D, "attention: head_dim must be 64 (D=%d, H=%d)", a.D, a.H);
  PEVIT_REQUIRE(a.r >= 0 && a.r <= 64, "attention: low-rank width r=%d out of range", a.r);
  return 0;
}

}  // namespace

int attn_delta_fwd_ref(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, const bf16* T,
                       const float* Qmat, const float* bias, bf16* o_tok, float* lse) {
  if (check_shape(a) != 0) return -1;
  const size_t smem = (2 * a.L * QS + AT_WARPS * a.L) * sizeof(float) + static_cast<size_t>(a.L) * KS * sizeof(bf16);
  PEVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_ref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(s, PC_ATTN_FWD);
  attn_fwd_ref_kernel<<<a.NB * a.H, AT_THREADS, smem, s>>>(a, q, k, v, T, Qmat, bias, o_tok, lse);
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int attn_delta_bwd_ref(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, const bf16* T,
                       const float* Qmat, const float* bias, const bf16* o_tok, const bf16* do_tok, const float* lse,
                       bf16* dqkv, int ld_dqkv, bf16* ddelta) {
  if (check_shape(a) != 0) return -1;
  const size_t smem = (2 * a.L * QS + 2 * AT_WARPS * a.L + 2 * a.L) * sizeof(float) +
                      2 * static_cast<size_t>(a.L) * KS * sizeof(bf16);
  PEVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_ref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(s, PC_ATTN_BWD);
  attn_bwd_ref_kernel<<<a.NB * a.H, AT_THREADS, smem, s>>>(a, q, k, v, T, Qmat, bias, o_tok, do_tok, lse, dqkv,
                                                            ld_dqkv, ddelta);
  PEVIT_CHECK_LAUNCH();
  return 0;
}

}  // namespace pevit
