// Small-factor kernels of the parameter-efficient updates.
//
// KAdaptation (reference evaluation/model.py:563-584) builds H = sum_i kron(u_i v_i^T, s_i t_i^T)
// as a materialised (32, D, D) einsum and multiplies x by it.  Here H is never formed:
// H = P Q^T with P[:, i] = u_i (x) s_i, Q[:, i] = v_i (x) t_i, so the
// forward only needs P^T appended as 2*32 extra rows of the in-projection weight (the QKV GEMM
// then emits T = X P for free) and Q for the in-attention expansion.  The factor gradients are
// the matching contractions of dP = X^T dT and dQ = alpha * dDelta^T T.
// LoRA (lora_model.py:490-514) is the same with P = A^T, Q = B, r = 4.
#include "common.cuh"
#include "kernels.h"

namespace pevit {
namespace {

__global__ void kad_expand_kernel(const float* __restrict__ u1, const float* __restrict__ v1,
                                  const float* __restrict__ u2, const float* __restrict__ v2,
                                  const float* __restrict__ sf, const float* __restrict__ tf, int D, float alpha,
                                  bf16* __restrict__ w_ext, bf16* __restrict__ w_ext_t, float* __restrict__ qmat,
                                  bf16* __restrict__ qmat_t, bf16* __restrict__ delta_w) {
  pdl_launch_dependents();
  pdl_wait();
  const int F = D / 32;
  const int ld_t = 3 * D + 64;
  const int total = 32 * D;  // (i, a, k)
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int i = idx / D, col = idx % D;  // col = a*F + k  (or c*F + p)
    const int a = col / F, k = col % F;
    const float s = sf[i * F + k], t = tf[i * F + k];
    const float pq = u1[i * 32 + a] * s, pv = u2[i * 32 + a] * s;
    const float qq = v1[i * 32 + a] * t, qv = v2[i * 32 + a] * t;
    w_ext[static_cast<size_t>(3 * D + i) * D + col] = __float2bfloat16(pq);
    w_ext[static_cast<size_t>(3 * D + 32 + i) * D + col] = __float2bfloat16(pv);
    w_ext_t[static_cast<size_t>(col) * ld_t + 3 * D + i] = __float2bfloat16(pq);
    w_ext_t[static_cast<size_t>(col) * ld_t + 3 * D + 32 + i] = __float2bfloat16(pv);
    qmat[static_cast<size_t>(col) * 32 + i] = qq;
    qmat[static_cast<size_t>(D + col) * 32 + i] = qv;
    qmat_t[static_cast<size_t>(i) * D + col] = __float2bfloat16(alpha * qq);
    qmat_t[static_cast<size_t>(32 + i) * D + col] = __float2bfloat16(alpha * qv);
    // delta_w[which][col][64]: q uses the T_q half of K, v the T_v half (the other half multiplies zeros)
    delta_w[static_cast<size_t>(col) * 64 + i] = __float2bfloat16(alpha * qq);
    delta_w[static_cast<size_t>(col) * 64 + 32 + i] = __float2bfloat16(0.f);
    delta_w[static_cast<size_t>(D + col) * 64 + i] = __float2bfloat16(0.f);
    delta_w[static_cast<size_t>(D + col) * 64 + 32 + i] = __float2bfloat16(alpha * qv);
  }
}

__global__ void lora_expand_kernel(const float* __restrict__ Aq, const float* __restrict__ Av,
                                   const float* __restrict__ Bq, const float* __restrict__ Bv, int D, int r,
                                   float alpha, bf16* __restrict__ w_ext, bf16* __restrict__ w_ext_t,
                                   float* __restrict__ qmat, bf16* __restrict__ qmat_t, bf16* __restrict__ delta_w) {
  pdl_launch_dependents();
  pdl_wait();
  const int ld_t = 3 * D + 2 * r;
  const int total = r * D;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int i = idx / D, col = idx % D;
    const float pq = Aq[i * D + col], pv = Av[i * D + col];
    const float qq = Bq[col * r + i], qv = Bv[col * r + i];
    w_ext[static_cast<size_t>(3 * D + i) * D + col] = __float2bfloat16(pq);
    w_ext[static_cast<size_t>(3 * D + r + i) * D + col] = __float2bfloat16(pv);
    w_ext_t[static_cast<size_t>(col) * ld_t + 3 * D + i] = __float2bfloat16(pq);
    w_ext_t[static_cast<size_t>(col) * ld_t + 3 * D + r + i] = __float2bfloat16(pv);
    qmat[static_cast<size_t>(col) * r + i] = qq;
    qmat[static_cast<size_t>(D + col) * r + i] = qv;
    qmat_t[static_cast<size_t>(i) * D + col] = __float2bfloat16(alpha * qq);
    qmat_t[static_cast<size_t>(r + i) * D + col] = __float2bfloat16(alpha * qv);
    delta_w[static_cast<size_t>(col) * 2 * r + i] = __float2bfloat16(alpha * qq);
    delta_w[static_cast<size_t>(col) * 2 * r + r + i] = __float2bfloat16(0.f);
    delta_w[static_cast<size_t>(D + col) * 2 * r + i] = __float2bfloat16(0.f);
    delta_w[static_cast<size_t>(D + col) * 2 * r + r + i] = __float2bfloat16(alpha * qv);
  }
}

// C[kc][nc] += scale * sum_m A[m][kc] * B[m][nc].  CUDA-core split-M version: each CTA owns a
// 64-wide slab of kc and a slice of rows, stages A (64 cols) and B (<= 64 cols) tiles of 32 rows in
// shared memory, then adds its partial with one atomic per output.
constexpr int ATB_TK = 64, ATB_TM = 32, ATB_THREADS = 256;

template <typename TA, typename TB>
__global__ void __launch_bounds__(ATB_THREADS)
atb_kernel(const TA* __restrict__ A, int lda, const TB* __restrict__ B, int ldb, int M, int Kc, int Nc, float scale,
           float* __restrict__ C, int rows_per_cta) {
  __shared__ float sA[ATB_TM][ATB_TK + 1];
  __shared__ float sB[ATB_TM][64 + 1];
  const int kc0 = blockIdx.x * ATB_TK;
  const int m_begin = blockIdx.y * rows_per_cta;
  const int m_end = min(M, m_begin + rows_per_cta);
  // thread owns outputs (kc = tid/4 .. , nc = (tid%4)*16 .. +16): 64 x 64 tile / 256 threads = 16 each
  const int tk = threadIdx.x >> 2;        // 0..63
  const int tn0 = (threadIdx.x & 3) * 16;  // 0,16,32,48
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  for (int m0 = m_begin; m0 < m_end; m0 += ATB_TM) {
    for (int idx = threadIdx.x; idx < ATB_TM * ATB_TK; idx += ATB_THREADS) {
      const int mm = idx / ATB_TK, kk = idx % ATB_TK;
      const int m = m0 + mm, kc = kc0 + kk;
      sA[mm][kk] = (m < m_end && kc < Kc) ? static_cast<float>(A[static_cast<size_t>(m) * lda + kc]) : 0.f;
    }
    for (int idx = threadIdx.x; idx < ATB_TM * 64; idx += ATB_THREADS) {
      const int mm = idx / 64, nn = idx % 64;
      const int m = m0 + mm;
      sB[mm][nn] = (m < m_end && nn < Nc) ? static_cast<float>(B[static_cast<size_t>(m) * ldb + nn]) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int mm = 0; mm < ATB_TM; ++mm) {
      const float av = sA[mm][tk];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = fmaf(av, sB[mm][tn0 + j], acc[j]);
    }
    __syncthreads();
  }
  const int kc = kc0 + tk;
  if (kc < Kc) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (tn0 + j < Nc) atomicAdd(C + static_cast<size_t>(kc) * Nc + tn0 + j, scale * acc[j]);
  }
}

// out[c] += sum_m X0[m][c] (+ X1[m][c]).  Thread = 8 columns (one 16-byte load) x a strided slice of rows;
// the 8 row-lanes of a block are reduced through shared memory, then one atomic per column per block.
constexpr int CS_COLG = 32, CS_ROWL = 8;  // 32 column groups x 8 row lanes = 256 threads
__global__ void __launch_bounds__(CS_COLG* CS_ROWL)
colsum_bf16_kernel(const bf16* __restrict__ X0, const bf16* __restrict__ X1, int ld, int M, int D,
                   float* __restrict__ out, int rows_per_cta) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[CS_ROWL][CS_COLG * 8 + 1];
  const int cg = threadIdx.x % CS_COLG, rl = threadIdx.x / CS_COLG;
  const int c0 = (blockIdx.x * CS_COLG + cg) * 8;
  const int m_begin = blockIdx.y * rows_per_cta;
  const int m_end = min(M, m_begin + rows_per_cta);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c0 + 8 <= D) {
    for (int m = m_begin + rl; m < m_end; m += CS_ROWL) {
      const uint4 a = *reinterpret_cast<const uint4*>(X0 + static_cast<size_t>(m) * ld + c0);
      const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) { const float2 f = unpack_bf16(w[t]); acc[2 * t] += f.x; acc[2 * t + 1] += f.y; }
      if (X1 != nullptr) {
        const uint4 b = *reinterpret_cast<const uint4*>(X1 + static_cast<size_t>(m) * ld + c0);
        const uint32_t u[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) { const float2 f = unpack_bf16(u[t]); acc[2 * t] += f.x; acc[2 * t + 1] += f.y; }
      }
    }
  } else {
    for (int j = 0; j < 8; ++j)
      if (c0 + j < D)
        for (int m = m_begin + rl; m < m_end; m += CS_ROWL) {
          acc[j] += __bfloat162float(X0[static_cast<size_t>(m) * ld + c0 + j]);
          if (X1 != nullptr) acc[j] += __bfloat162float(X1[static_cast<size_t>(m) * ld + c0 + j]);
        }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cg * 8 + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < CS_COLG * 8; c += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < CS_ROWL; ++r) t += red[r][c];
    const int col = blockIdx.x * CS_COLG * 8 + c;
    if (col < D) atomicAdd(out + col, t);
  }
}

// one CTA per Kronecker term i.  dP [D][64] (q | v), dQ [2][D][32].  Column i of dP / dQ is staged in shared
// memory first (all loads independent), then the 4*32 + 2*F small dot products run out of shared memory.
__global__ void __launch_bounds__(256)
kad_factor_grads_kernel(const float* __restrict__ dP, const float* __restrict__ dQ, const float* __restrict__ u1,
                        const float* __restrict__ v1, const float* __restrict__ u2, const float* __restrict__ v2,
                        const float* __restrict__ sf, const float* __restrict__ tf, int D, float* __restrict__ du1,
                        float* __restrict__ dv1, float* __restrict__ du2, float* __restrict__ dv2,
                        float* __restrict__ dsf, float* __restrict__ dtf, int accumulate) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];  // [4][D]: dPq, dPv, dQq, dQv columns i; then s, t, u1, u2, v1, v2 rows i
  const int i = blockIdx.x;
  const int F = D / 32;
  float* cPq = sm; float* cPv = sm + D; float* cQq = sm + 2 * D; float* cQv = sm + 3 * D;
  float* ss = sm + 4 * D; float* tt = ss + F; float* a1 = tt + F; float* a2 = a1 + 32; float* b1 = a2 + 32; float* b2 = b1 + 32;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    cPq[c] = dP[static_cast<size_t>(c) * 64 + i];
    cPv[c] = dP[static_cast<size_t>(c) * 64 + 32 + i];
    cQq[c] = dQ[static_cast<size_t>(c) * 32 + i];
    cQv[c] = dQ[static_cast<size_t>(D + c) * 32 + i];
  }
  for (int k = threadIdx.x; k < F; k += blockDim.x) { ss[k] = sf[i * F + k]; tt[k] = tf[i * F + k]; }
  if (threadIdx.x < 32) {
    a1[threadIdx.x] = u1[i * 32 + threadIdx.x]; a2[threadIdx.x] = u2[i * 32 + threadIdx.x];
    b1[threadIdx.x] = v1[i * 32 + threadIdx.x]; b2[threadIdx.x] = v2[i * 32 + threadIdx.x];
  }
  __syncthreads();
  // outputs 0..127: du1, du2, dv1, dv2 [a]; 128..128+2F: ds[k], dt[k]   (s and t are shared by q and v -- F2)
  for (int o = threadIdx.x; o < 128 + 2 * F; o += blockDim.x) {
    float g = 0.f;
    if (o < 128) {
      const int which = o >> 5, a = o & 31;
      const float* col = which == 0 ? cPq : (which == 1 ? cPv : (which == 2 ? cQq : cQv));
      const float* fac = which < 2 ? ss : tt;
      for (int k = 0; k < F; ++k) g = fmaf(col[a * F + k], fac[k], g);
      float* dst = which == 0 ? du1 : (which == 1 ? du2 : (which == 2 ? dv1 : dv2));
      dst[i * 32 + a] = accumulate ? dst[i * 32 + a] + g : g;
    } else if (o < 128 + F) {
      const int k = o - 128;
      for (int a = 0; a < 32; ++a) g = fmaf(cPq[a * F + k], a1[a], fmaf(cPv[a * F + k], a2[a], g));
      dsf[i * F + k] = accumulate ? dsf[i * F + k] + g : g;
    } else {
      const int k = o - 128 - F;
      for (int a = 0; a < 32; ++a) g = fmaf(cQq[a * F + k], b1[a], fmaf(cQv[a * F + k], b2[a], g));
      dtf[i * F + k] = accumulate ? dtf[i * F + k] + g : g;
    }
  }
}

// The same contractions for ALL layers of a backward pass in one launch (blockIdx.y = layer), added onto the callers'
// gradient buffers: ds / dt are per layer (plain +=, each element has one writer), du1 / dv1 / du2 / dv2 belong to the rule
// tensors every layer shares (model.py:1003-1010), so the layers' contributions meet there through atomicAdd.
struct KadLayers {
  const float* dP[KAD_MAX_LAYERS]; const float* dQ[KAD_MAX_LAYERS];
  const float* sf[KAD_MAX_LAYERS]; const float* tf[KAD_MAX_LAYERS];
  float* dsf[KAD_MAX_LAYERS]; float* dtf[KAD_MAX_LAYERS];
};
__global__ void __launch_bounds__(256)
kad_factor_grads_batch_kernel(const __grid_constant__ KadLayers ly, const float* __restrict__ u1, const float* __restrict__ v1,
                              const float* __restrict__ u2, const float* __restrict__ v2, int D, float* __restrict__ du1,
                              float* __restrict__ dv1, float* __restrict__ du2, float* __restrict__ dv2) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];
  const int i = blockIdx.x, layer = blockIdx.y;
  const int F = D / 32;
  const float* dP = ly.dP[layer]; const float* dQ = ly.dQ[layer];
  const float* sf = ly.sf[layer]; const float* tf = ly.tf[layer];
  float* dsf = ly.dsf[layer]; float* dtf = ly.dtf[layer];
  float* cPq = sm; float* cPv = sm + D; float* cQq = sm + 2 * D; float* cQv = sm + 3 * D;
  float* ss = sm + 4 * D; float* tt = ss + F; float* a1 = tt + F; float* a2 = a1 + 32; float* b1 = a2 + 32; float* b2 = b1 + 32;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    cPq[c] = dP[static_cast<size_t>(c) * 64 + i];
    cPv[c] = dP[static_cast<size_t>(c) * 64 + 32 + i];
    cQq[c] = dQ[static_cast<size_t>(c) * 32 + i];
    cQv[c] = dQ[static_cast<size_t>(D + c) * 32 + i];
  }
  for (int k = threadIdx.x; k < F; k += blockDim.x) { ss[k] = sf[i * F + k]; tt[k] = tf[i * F + k]; }
  if (threadIdx.x < 32) {
    a1[threadIdx.x] = u1[i * 32 + threadIdx.x]; a2[threadIdx.x] = u2[i * 32 + threadIdx.x];
    b1[threadIdx.x] = v1[i * 32 + threadIdx.x]; b2[threadIdx.x] = v2[i * 32 + threadIdx.x];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 128 + 2 * F; o += blockDim.x) {
    float g = 0.f;
    if (o < 128) {
      const int which = o >> 5, a = o & 31;
      const float* col = which == 0 ? cPq : (which == 1 ? cPv : (which == 2 ? cQq : cQv));
      const float* fac = which < 2 ? ss : tt;
      for (int k = 0; k < F; ++k) g = fmaf(col[a * F + k], fac[k], g);
      float* dst = which == 0 ? du1 : (which == 1 ? du2 : (which == 2 ? dv1 : dv2));
      atomicAdd(dst + i * 32 + a, g);
    } else if (o < 128 + F) {
      const int k = o - 128;
      for (int a = 0; a < 32; ++a) g = fmaf(cPq[a * F + k], a1[a], fmaf(cPv[a * F + k], a2[a], g));
      dsf[i * F + k] += g;
    } else {
      const int k = o - 128 - F;
      for (int a = 0; a < 32; ++a) g = fmaf(cQq[a * F + k], b1[a], fmaf(cQv[a * F + k], b2[a], g));
      dtf[i * F + k] += g;
    }
  }
}

__global__ void cast2d_kernel(const float* __restrict__ src, int lds, bf16* __restrict__ dst, int ldd, int rows,
                              int cols) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t total = static_cast<size_t>(rows) * cols;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = idx / cols, c = idx % cols;
    dst[r * ldd + c] = __float2bfloat16(src[r * lds + c]);
  }
}

__global__ void cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

__global__ void transpose_cast_kernel(const float* __restrict__ src, int rows, int cols, bf16* __restrict__ dst,
                                      int ldd) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? src[static_cast<size_t>(r) * cols + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < rows) dst[static_cast<size_t>(c) * ldd + r] = __float2bfloat16(tile[threadIdx.x][j]);
  }
}

int grid_for(size_t n, int threads) {
  size_t g = (n + threads - 1) / threads;
  const size_t cap = static_cast<size_t>(sm_count()) * 8;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int kad_expand(cudaStream_t s, const float* u1, const float* v1, const float* u2, const float* v2, const float* sfac,
               const float* tfac, int D, float alpha, bf16* w_ext, bf16* w_ext_t, float* qmat, bf16* qmat_t,
               bf16* delta_w) {
  PEVIT_REQUIRE(D % 32 == 0, "kad_expand: D=%d not divisible by phm_dim 32", D);
  ProfScope prof(s, PC_EXPAND);
  PEVIT_CHECK_CUDA(launch_kernel(kad_expand_kernel, dim3(grid_for(32 * static_cast<size_t>(D), 256)), dim3(256), 0, s, 1, u1, v1,
                                 u2, v2, sfac, tfac, D, alpha, w_ext, w_ext_t, qmat, qmat_t, delta_w));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int lora_expand(cudaStream_t s, const float* Aq, const float* Av, const float* Bq, const float* Bv, int D, int r,
                float alpha, bf16* w_ext, bf16* w_ext_t, float* qmat, bf16* qmat_t, bf16* delta_w) {
  ProfScope prof(s, PC_EXPAND);
  PEVIT_CHECK_CUDA(launch_kernel(lora_expand_kernel, dim3(grid_for(static_cast<size_t>(r) * D, 256)), dim3(256), 0, s, 1, Aq, Av,
                                 Bq, Bv, D, r, alpha, w_ext, w_ext_t, qmat, qmat_t, delta_w));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int atb_accumulate(cudaStream_t s, const void* A, int a_is_bf16, int lda, const void* B, int b_is_bf16, int ldb, int M,
                   int Kc, int Nc, float scale, float* C) {
  PEVIT_REQUIRE(Nc >= 1 && Nc <= 64, "atb_accumulate: Nc=%d must be in 1..64", Nc);
  const int gx = (Kc + ATB_TK - 1) / ATB_TK;
  int splits = (sm_count() * 4 + gx - 1) / gx;
  int rows_per_cta = (M + splits - 1) / splits;
  rows_per_cta = ((rows_per_cta + ATB_TM - 1) / ATB_TM) * ATB_TM;
  splits = (M + rows_per_cta - 1) / rows_per_cta;
  dim3 grid(gx, splits);
  ProfScope prof(s, PC_ATB);
  if (a_is_bf16 && b_is_bf16)
    atb_kernel<bf16, bf16><<<grid, ATB_THREADS, 0, s>>>((const bf16*)A, lda, (const bf16*)B, ldb, M, Kc, Nc, scale, C, rows_per_cta);
  else if (a_is_bf16)
    atb_kernel<bf16, float><<<grid, ATB_THREADS, 0, s>>>((const bf16*)A, lda, (const float*)B, ldb, M, Kc, Nc, scale, C, rows_per_cta);
  else if (b_is_bf16)
    atb_kernel<float, bf16><<<grid, ATB_THREADS, 0, s>>>((const float*)A, lda, (const bf16*)B, ldb, M, Kc, Nc, scale, C, rows_per_cta);
  else
    atb_kernel<float, float><<<grid, ATB_THREADS, 0, s>>>((const float*)A, lda, (const float*)B, ldb, M, Kc, Nc, scale, C, rows_per_cta);
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int colsum_bf16(cudaStream_t s, const bf16* X0, const bf16* X1, int ld, int M, int D, float* out) {
  PEVIT_REQUIRE(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(X0) & 15) == 0 && (reinterpret_cast<uintptr_t>(X1) & 15) == 0,
                "colsum_bf16: rows must be 16-byte aligned (ld=%d)", ld);
  const int gx = (D + CS_COLG * 8 - 1) / (CS_COLG * 8);
  int splits = (sm_count() * 2 + gx - 1) / gx;
  int rows_per_cta = (M + splits - 1) / splits;
  if (rows_per_cta < CS_ROWL) rows_per_cta = CS_ROWL;
  splits = (M + rows_per_cta - 1) / rows_per_cta;
  ProfScope prof(s, PC_COLSUM);
  PEVIT_CHECK_CUDA(launch_kernel(colsum_bf16_kernel, dim3(gx, splits), dim3(CS_COLG * CS_ROWL), 0, s, 1, X0, X1, ld, M, D, out,
                                 rows_per_cta));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int kad_factor_grads(cudaStream_t s, const float* dP, const float* dQ, const float* u1, const float* v1, const float* u2,
                     const float* v2, const float* sfac, const float* tfac, int D, float* du1, float* dv1, float* du2,
                     float* dv2, float* dsfac, float* dtfac, bool accumulate) {
  ProfScope prof(s, PC_FACTOR_GRADS);
  const size_t smem = (4 * static_cast<size_t>(D) + 2 * (D / 32) + 128) * sizeof(float);
  PEVIT_CHECK_CUDA(launch_kernel(kad_factor_grads_kernel, dim3(32), dim3(256), smem, s, 1, dP, dQ, u1, v1, u2, v2, sfac, tfac, D,
                                 du1, dv1, du2, dv2, dsfac, dtfac, accumulate ? 1 : 0));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int kad_factor_grads_batch(cudaStream_t s, int count, const float* const* dP, const float* const* dQ,
                           const float* const* sfac, const float* const* tfac, float* const* dsfac, float* const* dtfac,
                           const float* u1, const float* v1, const float* u2, const float* v2, int D, float* du1, float* dv1,
                           float* du2, float* dv2) {
  PEVIT_REQUIRE(count >= 1 && count <= KAD_MAX_LAYERS, "kad_factor_grads_batch: %d layers (1..%d)", count, KAD_MAX_LAYERS);
  PEVIT_REQUIRE(D % 32 == 0, "kad_factor_grads_batch: D=%d not divisible by phm_dim 32", D);
  KadLayers ly{};
  for (int l = 0; l < count; ++l) {
    PEVIT_REQUIRE(dP[l] && dQ[l] && sfac[l] && tfac[l] && dsfac[l] && dtfac[l], "kad_factor_grads_batch: null pointer in layer %d", l);
    ly.dP[l] = dP[l]; ly.dQ[l] = dQ[l]; ly.sf[l] = sfac[l]; ly.tf[l] = tfac[l]; ly.dsf[l] = dsfac[l]; ly.dtf[l] = dtfac[l];
  }
  ProfScope prof(s, PC_FACTOR_GRADS);
  const size_t smem = (4 * static_cast<size_t>(D) + 2 * (D / 32) + 128) * sizeof(float);
  PEVIT_CHECK_CUDA(launch_kernel(kad_factor_grads_batch_kernel, dim3(32, count), dim3(256), smem, s, 1, ly, u1, v1, u2, v2, D,
                                 du1, dv1, du2, dv2));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int cast_f32_to_bf16_2d(cudaStream_t s, const float* src, int lds, bf16* dst, int ldd, int rows, int cols) {
  ProfScope prof(s, PC_CAST);
  PEVIT_CHECK_CUDA(launch_kernel(cast2d_kernel, dim3(grid_for(static_cast<size_t>(rows) * cols, 256)), dim3(256), 0, s, 1, src,
                                 lds, dst, ldd, rows, cols));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int cast_f32_to_bf16(cudaStream_t s, const float* src, bf16* dst, size_t n) {
  ProfScope prof(s, PC_CAST);
  PEVIT_CHECK_CUDA(launch_kernel(cast_kernel, dim3(grid_for(n, 256)), dim3(256), 0, s, 1, src, dst, n));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int transpose_f32_to_bf16(cudaStream_t s, const float* src, int rows, int cols, bf16* dst, int ldd) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  ProfScope prof(s, PC_CAST);
  transpose_cast_kernel<<<grid, block, 0, s>>>(src, rows, cols, dst, ldd);
  PEVIT_CHECK_LAUNCH();
  return 0;
}

}  // namespace pevit
