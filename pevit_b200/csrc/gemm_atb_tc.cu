// C[kc][n] += scale * sum_m A[m][kc] * B[m][n]  -- the "weight-gradient shaped" products of the PEFT
// tensors (dP = X^T dT, dQ = alpha dDelta^T T, dW_up = dY^T u, dW_down^T = a_n^T dzd): the contraction
// runs over the L*N token rows, the outputs are tiny (D x <=64).  tcgen05 with BOTH operands MN-major
// (row-major [m][..] tiles as TMA delivers them, no transposes), split over the token rows across
// CTAs, fp32 partials reduced with red.global.add.  Replaces the autograd MulBackward/SumBackward over
// the reference's materialised Kronecker einsum (evaluation/model.py:406-417).
#include "common.cuh"
#include "kernels.h"

namespace pevit {
namespace {

constexpr int ATB_BK = 64;                       // token rows per pipeline stage
constexpr int ATB_STAGES = 4;
constexpr int ATB_BLK = 64 * 128;                // one [64 rows][64 cols] bf16 block
constexpr int ATB_STAGE_BYTES = 3 * ATB_BLK;     // A: two 64-wide kc blocks, B: one 64-wide n block
constexpr int ATB_SMEM = ATB_STAGES * ATB_STAGE_BYTES + 128 + 1024;
constexpr int ATB_THREADS = 192;

struct AtbParams {
  int M, Kc, n_lo, n_cnt, ldc, rows_per_split, vec4;
  float scale;
  float* C;
};

constexpr int ATB_MAX_BATCH = 3;
constexpr int ATB_CS_SLICES = 2;  // z-slices of the grid that run the column-sum job
struct AtbBatch {
  CUtensorMap tm_a[ATB_MAX_BATCH], tm_b[ATB_MAX_BATCH];
  AtbParams p[ATB_MAX_BATCH];
  int count;
  AtbColsum cs;  // optional rider: out[c] += sum_m X0[m][c] + X1[m][c]  (cs.out == nullptr: none)
};

// The column-sum rider (KAdaptation's shared bias gradient, colsum(dDelta_q) + colsum(dDelta_v), model.py:583): the same
// d(delta) planes the products above read, summed by the CTAs of the extra z-slices while the products run -- one launch
// and one pass through L2 instead of a second kernel behind this one.  192 threads = (D / 8 column groups of one
// 16-byte load) x row lanes; a CTA owns a contiguous slice of rows; row lanes meet in shared memory, then one atomic per
// column per CTA.
__device__ __forceinline__ void colsum_rider(const AtbColsum& cs, float* red, int cta, int nctas) {
  const int ncg = cs.D >> 3;                      // host guarantees D % 8 == 0 and ncg <= ATB_THREADS
  const int lanes = ATB_THREADS / ncg;            // row lanes per CTA
  const int cg = threadIdx.x % ncg, rl = threadIdx.x / ncg;
  const int rows_per = (cs.M + nctas - 1) / nctas;
  const int m_begin = cta * rows_per, m_end = min(cs.M, m_begin + rows_per);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (rl < lanes) {
#pragma unroll 4
    for (int m = m_begin + rl; m < m_end; m += lanes) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(cs.X0 + static_cast<size_t>(m) * cs.ld) + cg);
      const uint4 b = cs.X1 != nullptr ? __ldg(reinterpret_cast<const uint4*>(cs.X1 + static_cast<size_t>(m) * cs.ld) + cg)
                                       : make_uint4(0, 0, 0, 0);
      const uint32_t w[4] = {a.x, a.y, a.z, a.w}, u[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = unpack_bf16(w[t]), h = unpack_bf16(u[t]);
        acc[2 * t] += f.x + h.x;
        acc[2 * t + 1] += f.y + h.y;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[rl * cs.D + cg * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cs.D; c += ATB_THREADS) {
    float t = 0.f;
    for (int r = 0; r < lanes; ++r) t += red[r * cs.D + c];
    atomicAdd(cs.out + c, t);
  }
}

// blockIdx.z selects one of up to three independent products (same M): the block backward has three of them per
// layer (dQ_q, dQ_v, dP), each far too short to fill the GPU or amortise a launch on its own.
__global__ void __launch_bounds__(ATB_THREADS)
atb_tc_kernel(const __grid_constant__ AtbBatch batch) {
  if (static_cast<int>(blockIdx.z) >= batch.count) {   // column-sum rider slices (whole CTAs: no barrier is shared)
    extern __shared__ uint8_t smem_cs[];
    pdl_launch_dependents();
    pdl_wait();
    const int per_slice = gridDim.x * gridDim.y;
    colsum_rider(batch.cs, reinterpret_cast<float*>(smem_cs),
                 (static_cast<int>(blockIdx.z) - batch.count) * per_slice + blockIdx.y * gridDim.x + blockIdx.x,
                 per_slice * ATB_CS_SLICES);
    return;
  }
  const CUtensorMap& tm_a = batch.tm_a[blockIdx.z];
  const CUtensorMap& tm_b = batch.tm_b[blockIdx.z];
  const AtbParams& p = batch.p[blockIdx.z];
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + ATB_STAGES * ATB_STAGE_BYTES);
  uint64_t* empty = full + ATB_STAGES;
  uint64_t* done = empty + ATB_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kc0 = blockIdx.x * 128;
  const int m_begin = blockIdx.y * p.rows_per_split;
  const int m_end = min(p.M, m_begin + p.rows_per_split);
  const int num_kb = (m_end - m_begin + ATB_BK - 1) / ATB_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_b);
    for (int s = 0; s < ATB_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % ATB_STAGES;
        mbar_wait(&empty[s], ((kb / ATB_STAGES) & 1) ^ 1);
        uint8_t* st = smem + s * ATB_STAGE_BYTES;
        const int m0 = m_begin + kb * ATB_BK;
        mbar_expect_tx(&full[s], ATB_STAGE_BYTES);
        tma_load_2d(st, &tm_a, &full[s], kc0, m0);
        tma_load_2d(st + ATB_BLK, &tm_a, &full[s], kc0 + 64, m0);
        tma_load_2d(st + 2 * ATB_BLK, &tm_b, &full[s], 0, m0);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64) | IDESC_A_MN | IDESC_B_MN;
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % ATB_STAGES;
      mbar_wait(&full[s], (kb / ATB_STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = smem_u32(smem + s * ATB_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < ATB_BK / 16; ++k)
          umma_bf16_ss(tmem_base, umma_desc_mnmajor_sw128(st + k * 2048, ATB_BLK),
                       umma_desc_mnmajor_sw128(st + 2 * ATB_BLK + k * 2048, ATB_BLK), idesc, (kb | k) != 0);
        umma_commit(&empty[s]);
        if (kb == num_kb - 1) umma_commit(done);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int kc = kc0 + quad * 32 + lane;
    if (num_kb > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c, v);
        tmem_ld_wait();
        if (kc < p.Kc) {
          float* crow = p.C + static_cast<size_t>(kc) * p.ldc;
          if (p.vec4) {  // 16-byte vector reductions: window and row stride are multiples of 4 floats
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int n = c + j - p.n_lo;
              if (n >= 0 && n < p.n_cnt)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow + n),
                             "f"(p.scale * __uint_as_float(v[j])), "f"(p.scale * __uint_as_float(v[j + 1])),
                             "f"(p.scale * __uint_as_float(v[j + 2])), "f"(p.scale * __uint_as_float(v[j + 3]))
                             : "memory");
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = c + j - p.n_lo;
              if (n >= 0 && n < p.n_cnt) atomicAdd(crow + n, p.scale * __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace

// A: bf16 [M][lda] (uses Kc columns).  B: bf16 [M][ldb] with nb_cols (<= 64) columns visible; columns
// [n_lo, n_lo + n_cnt) of the product are accumulated into C[kc][0 .. n_cnt) (row stride ldc).
bool atb_colsum_rider_supported(int D, int ld) { return D % 8 == 0 && D / 8 <= ATB_THREADS && D >= 8 && ld % 8 == 0; }

int atb_tc_batch(cudaStream_t s, const AtbProblem* probs, int count, int M, int Kc, const AtbColsum* colsum) {
  PEVIT_REQUIRE(count >= 1 && count <= ATB_MAX_BATCH, "atb_tc_batch: %d problems (1..%d)", count, ATB_MAX_BATCH);
  if (colsum != nullptr)
    PEVIT_REQUIRE(colsum->X0 != nullptr && colsum->out != nullptr && atb_colsum_rider_supported(colsum->D, colsum->ld) &&
                      (reinterpret_cast<uintptr_t>(colsum->X0) & 15) == 0 && (reinterpret_cast<uintptr_t>(colsum->X1) & 15) == 0 &&
                      static_cast<size_t>(ATB_THREADS / (colsum->D / 8)) * colsum->D * sizeof(float) <= ATB_SMEM,
                  "atb_tc_batch: column-sum rider needs 16-byte rows and D / 8 <= %d (D=%d ld=%d)", ATB_THREADS, colsum->D,
                  colsum->ld);
  const int gx = (Kc + 127) / 128;
  int splits = (sm_count() + gx * count - 1) / (gx * count);  // ~one CTA per SM over the whole batch
  if (splits < 8) splits = 8;
  int rps = (M + splits - 1) / splits;
  rps = ((rps + ATB_BK - 1) / ATB_BK) * ATB_BK;
  splits = (M + rps - 1) / rps;
  AtbBatch batch;
  batch.count = count;
  batch.cs = colsum != nullptr ? *colsum : AtbColsum{nullptr, nullptr, 0, 0, 0, nullptr};
  for (int i = 0; i < count; ++i) {
    const AtbProblem& q = probs[i];
    PEVIT_REQUIRE(q.nb_cols >= 1 && q.nb_cols <= 64 && q.n_lo >= 0 && q.n_lo + q.n_cnt <= q.nb_cols,
                  "atb_tc: column window [%d,%d) outside the %d visible columns", q.n_lo, q.n_lo + q.n_cnt, q.nb_cols);
    PEVIT_REQUIRE(q.lda % 8 == 0 && q.ldb % 8 == 0 && (reinterpret_cast<uintptr_t>(q.A) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(q.B) & 15) == 0,
                  "atb_tc: operands need 16-byte aligned rows (lda=%d ldb=%d)", q.lda, q.ldb);
    if (make_tmap_bf16_2d(&batch.tm_a[i], q.A, M, Kc, q.lda, ATB_BK, 64) != 0) return -1;
    if (make_tmap_bf16_2d(&batch.tm_b[i], q.B, M, q.nb_cols, q.ldb, ATB_BK, 64) != 0) return -1;
    const int vec4 = (q.n_lo % 4 == 0 && q.n_cnt % 4 == 0 && q.ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(q.C) & 15) == 0) ? 1 : 0;
    batch.p[i] = AtbParams{M, Kc, q.n_lo, q.n_cnt, q.ldc, rps, vec4, q.scale, q.C};
  }
  static bool configured[64] = {};
  int dev = 0;
  PEVIT_CHECK_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    PEVIT_CHECK_CUDA(cudaFuncSetAttribute(atb_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATB_SMEM));
    configured[dev & 63] = true;
  }
  ProfScope prof(s, PC_ATB);
  PEVIT_CHECK_CUDA(launch_kernel(atb_tc_kernel, dim3(gx, splits, count + (colsum != nullptr ? ATB_CS_SLICES : 0)),
                                 dim3(ATB_THREADS), ATB_SMEM, s, 1, batch));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int atb_tc(cudaStream_t s, const bf16* A, int lda, const bf16* B, int ldb, int nb_cols, int M, int Kc, int n_lo,
           int n_cnt, float scale, float* C, int ldc) {
  const AtbProblem q{A, lda, B, ldb, nb_cols, n_lo, n_cnt, scale, C, ldc};
  return atb_tc_batch(s, &q, 1, M, Kc, nullptr);
}

}  // namespace pevit
