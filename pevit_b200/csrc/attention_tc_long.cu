// Attention core on tcgen05 / TMEM for sequences longer than one 128-row tile (ViT-B/16: L = 197, ViT-L/14:
// L = 257; any 128 < L <= 384).  Same contract as attention_tc.cu (reference evaluation/model.py:803-815; q', k, v'
// head-major bf16 with the low-rank delta already applied), same building blocks -- 128 x 128 score tiles, TMA ring,
// accumulators in TMEM, one softmax thread per query row -- composed over (query tile t, key block j) pairs:
//
//   forward   work item = (image-head g, query tile t).  For every key block j: S_j = Q_t K_j^T, block-local
//             softmax statistics (m_j, l_j), O_j = exp(S_j - m_j) V_j in its own TMEM buffer; the item's epilogue
//             merges the blocks exactly: O = sum_j e^{m_j - m} O_j / sum_j e^{m_j - m} l_j.  Nothing is rescaled in
//             TMEM and no partial result touches HBM.
//   backward  work item = (g, group of <= 2 query tiles).  Key blocks outer, query tiles inner: dK_j, dV_j accumulate
//             in TMEM over the tiles, dQ_t over the key blocks (one TMEM buffer per tile of the group: 512 columns in
//             all).  P = exp(S - lse) uses the saved row statistic, delta = rowsum(dO o O) is recomputed from the O
//             tile that rides along in the TMA stage.  A second group of the same head (L > 256) adds its dK / dV
//             contribution onto the first group's output.
//
// Rows / keys beyond L inside a 128-row box belong to the next head (or are zero-filled past the end of the
// tensor): they are masked to exact zeros in P and dS, so they never contribute, and never stored.
#include "common.cuh"
#include "kernels.h"

namespace pevit {
namespace {

constexpr int TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16
constexpr float LOG2E = 1.4426950408889634f;
constexpr int MAX_BLOCKS = 3;          // key blocks / query tiles per head (L <= 384)

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ===================================================================================================== forward
constexpr int FL_STAGES = 3;
constexpr int FL_STAGE_BYTES = 3 * TILE_BYTES;  // Q_t, K_j, V_j
constexpr int FL_P_BYTES = 2 * TILE_BYTES;      // [128 rows][128 keys] bf16
constexpr int FL_STATS_BYTES = 2 * MAX_BLOCKS * 128 * 8;
constexpr int FL_SMEM = FL_STAGES * FL_STAGE_BYTES + 2 * FL_P_BYTES + FL_STATS_BYTES + 256 + 1024;
constexpr int FL_THREADS = 320;  // 2 softmax warpgroups, TMA warp, MMA warp

struct FwdLongParams {
  int L, NB, H, D, heads_total, nq, nkv, num_items;
  bf16* o_tok;
  float* lse;
};

__global__ void __launch_bounds__(FL_THREADS, 1)
attn_fwd_long_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, FwdLongParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sP = smem + FL_STAGES * FL_STAGE_BYTES;
  float2* stats = reinterpret_cast<float2*>(sP + 2 * FL_P_BYTES);  // [item parity][block j][row] = (max, sum)
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stats) + FL_STATS_BYTES);
  uint64_t* full = bars;                  // [FL_STAGES]
  uint64_t* empty = full + FL_STAGES;     // [FL_STAGES]
  uint64_t* s_full = empty + FL_STAGES;   // [2]  S_b ready (MMA -> softmax warpgroup b)
  uint64_t* p_full = s_full + 2;          // [2]  P_b written (128 arrivals)
  uint64_t* o_full = p_full + 2;          // all O_j of the item accumulated
  uint64_t* o_empty = o_full + 1;         // O buffers drained (8 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L, nkv = p.nkv;
  const int n_local = (p.num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                      static_cast<int>(gridDim.x);
  const int n_blocks = n_local * nkv;  // (item, key block) pairs this CTA walks through

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    for (int s = 0; s < FL_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&p_full[b], 128); }
    mbar_init(o_full, 1);
    mbar_init(o_empty, 8);
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  // TMEM columns: S0 [0,128) S1 [128,256) O_j [256 + 64 j, +64)

  if (warp == 8) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      for (int bi = 0; bi < n_blocks; ++bi) {
        const int k = bi / nkv, j = bi - k * nkv;
        const int item = blockIdx.x + k * gridDim.x;
        const int g = item / p.nq, t = item - g * p.nq;
        const int s = bi % FL_STAGES;
        mbar_wait(&empty[s], ((bi / FL_STAGES) & 1) ^ 1);
        uint8_t* st = smem + s * FL_STAGE_BYTES;
        mbar_expect_tx(&full[s], FL_STAGE_BYTES);
        tma_load_2d(st, &tm_q, &full[s], 0, g * L + t * 128);
        tma_load_2d(st + TILE_BYTES, &tm_k, &full[s], 0, g * L + j * 128);
        tma_load_2d(st + 2 * TILE_BYTES, &tm_v, &full[s], 0, g * L + j * 128);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64) | IDESC_B_MN;
    auto issue_s = [&](int bi) {
      const int s = bi % FL_STAGES, b = bi & 1;
      mbar_wait(&full[s], (bi / FL_STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sq = smem_u32(smem + s * FL_STAGE_BYTES);
        const uint64_t dq = umma_desc_kmajor_sw128(sq);
        const uint64_t dk = umma_desc_kmajor_sw128(sq + TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem_base + b * 128, dq + 2 * kk, dk + 2 * kk, idesc_s, kk != 0);
        umma_commit(&s_full[b]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int bi) {
      const int s = bi % FL_STAGES, b = bi & 1;
      const int k = bi / nkv, j = bi - k * nkv;
      mbar_wait(&p_full[b], (bi >> 1) & 1);
      if (j == 0) mbar_wait(o_empty, (k & 1) ^ 1);  // the previous item's epilogue has drained the O buffers
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sp = smem_u32(sP + b * FL_P_BYTES);
        const uint32_t sv = smem_u32(smem + s * FL_STAGE_BYTES + 2 * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t da = umma_desc_kmajor_sw128(sp + (kk >> 2) * TILE_BYTES) + 2 * (kk & 3);
          const uint64_t db = umma_desc_mnmajor_sw128(sv + kk * 2048, 8192);
          umma_bf16_ss(tmem_base + 256 + j * 64, da, db, idesc_o, kk != 0);
        }
        umma_commit(&empty[s]);
        if (j == nkv - 1) umma_commit(o_full);
      }
      __syncwarp();
    };
    if (n_blocks > 0) issue_s(0);
    for (int bi = 0; bi < n_blocks; ++bi) {
      if (bi + 1 < n_blocks) issue_s(bi + 1);
      issue_pv(bi);
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroups + item epilogue
    const int grp = warp >> 2, quad = warp & 3;
    const int row = quad * 32 + lane;
    uint8_t* myP = sP + grp * FL_P_BYTES;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    int cnt = 0;  // blocks this warpgroup has processed (S_grp / P_grp use count)
    for (int k = 0; k < n_local; ++k) {
      const int item = blockIdx.x + k * gridDim.x;
      const int g = item / p.nq, t = item - g * p.nq;
      float2* st_item = stats + (k & 1) * MAX_BLOCKS * 128;
      for (int j = 0; j < nkv; ++j) {
        const int bi = k * nkv + j;
        if ((bi & 1) != grp) continue;
        const int nvalid = min(128, L - j * 128);  // keys of this block that exist
        mbar_wait(&s_full[grp], cnt & 1);
        tc_fence_after();
        float sc[128];
        {
          uint32_t v[128];
#pragma unroll
          for (int c = 0; c < 128; c += 32) tmem_ld_32x32(t_lane + grp * 128 + c, *reinterpret_cast<uint32_t(*)[32]>(&v[c]));
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 128; ++jj) sc[jj] = __uint_as_float(v[jj]);
        }
        // four independent max / sum chains: with few warps per scheduler, ILP is the latency hiding
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int jj = 0; jj < 128; ++jj) if (jj < nvalid) m4[jj & 3] = fmaxf(m4[jj & 3], sc[jj]);
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        const float mxs = mx * LOG2E;
#pragma unroll
        for (int jj = 0; jj < 128; ++jj) {
          const float e = (jj < nvalid) ? fast_exp2(fmaf(sc[jj], LOG2E, -mxs)) : 0.f;
          sc[jj] = e;
          s4[jj & 3] += e;
        }
        const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const uint4 pk = make_uint4(pack_bf16(sc[8 * c], sc[8 * c + 1]), pack_bf16(sc[8 * c + 2], sc[8 * c + 3]),
                                      pack_bf16(sc[8 * c + 4], sc[8 * c + 5]), pack_bf16(sc[8 * c + 6], sc[8 * c + 7]));
          const int half = c >> 3, ch = c & 7;
          *reinterpret_cast<uint4*>(myP + half * TILE_BYTES + row * 128 + ((ch ^ (row & 7)) << 4)) = pk;
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&p_full[grp]);
        st_item[j * 128 + row] = make_float2(mx, sum);
        ++cnt;
      }
      // ---- item epilogue: both warpgroups, 32 of the 64 head columns each
      named_bar_sync(3, 256);  // every block's (max, sum) of this item is in shared memory
      float m = -INFINITY;
      for (int j = 0; j < nkv; ++j) m = fmaxf(m, st_item[j * 128 + row].x);
      float w[MAX_BLOCKS], l = 0.f;
#pragma unroll
      for (int j = 0; j < MAX_BLOCKS; ++j) {
        w[j] = 0.f;
        if (j < nkv) {
          const float2 s2 = st_item[j * 128 + row];
          w[j] = fast_exp2((s2.x - m) * LOG2E);
          l = fmaf(w[j], s2.y, l);
        }
      }
      const float inv = 1.f / l;
      mbar_wait(o_full, k & 1);
      tc_fence_after();
      float acc[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] = 0.f;
#pragma unroll
      for (int j = 0; j < MAX_BLOCKS; ++j) {
        if (j < nkv) {
          uint32_t v[32];
          tmem_ld_32x32(t_lane + 256 + j * 64 + grp * 32, v);
          tmem_ld_wait();
          const float wj = w[j] * inv;
#pragma unroll
          for (int c = 0; c < 32; ++c) acc[c] = fmaf(wj, __uint_as_float(v[c]), acc[c]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      const int lq = t * 128 + row;
      if (lq < L && g < p.heads_total) {
        const int n = g / p.H, h = g - n * p.H;
        bf16* orow = p.o_tok + (static_cast<size_t>(lq) * p.NB + n) * p.D + h * 64 + grp * 32;
#pragma unroll
        for (int c = 0; c < 32; c += 8)
          *reinterpret_cast<uint4*>(orow + c) = make_uint4(pack_bf16(acc[c], acc[c + 1]), pack_bf16(acc[c + 2], acc[c + 3]),
                                                           pack_bf16(acc[c + 4], acc[c + 5]), pack_bf16(acc[c + 6], acc[c + 7]));
        if (grp == 0) p.lse[static_cast<size_t>(g) * L + lq] = m + __logf(l);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ===================================================================================================== backward
constexpr int BL_STAGES = 2;
constexpr int BL_STAGE_BYTES = 5 * TILE_BYTES;  // Q_t, K_j, V_j, dO_t, O_t
constexpr int BL_P_BYTES = 2 * TILE_BYTES;
constexpr int BL_SMEM = BL_STAGES * BL_STAGE_BYTES + 2 * BL_P_BYTES + 256 + 1024;
constexpr int BL_THREADS = 320;  // WG0 (P, dS), WG1 (gradients out), TMA warp, MMA warp

struct BwdLongParams {
  int L, NB, H, D, heads_total, nq, nkv, ngroups, num_items, ld;
  const float* lse;
  bf16* dqkv;
  bf16* ddelta;  // nullable
};

__device__ __forceinline__ void add_bf16x8(float (&f)[8], const uint4& x) {
  const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 a = unpack_bf16(w[t]);
    f[2 * t] += a.x;
    f[2 * t + 1] += a.y;
  }
}

__global__ void __launch_bounds__(BL_THREADS, 1)
attn_bwd_long_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                     const __grid_constant__ CUtensorMap tm_o, BwdLongParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sP = smem + BL_STAGES * BL_STAGE_BYTES;
  uint8_t* sdS = sP + BL_P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + BL_P_BYTES);
  uint64_t* full = bars;                   // [BL_STAGES]
  uint64_t* empty = full + BL_STAGES;      // [BL_STAGES]
  uint64_t* s_full = empty + BL_STAGES;    // MMA1 (S, dP) of a pair done
  uint64_t* pds_full = s_full + 1;         // WG0 wrote P, dS (128 arrivals)
  uint64_t* mma2_done = pds_full + 1;      // MMA2 of a pair done reading P, dS
  uint64_t* kv_full = mma2_done + 1;       // dK_j, dV_j complete (all tiles of the group)
  uint64_t* kv_empty = kv_full + 1;        // WG1 drained dK_j, dV_j (128 arrivals)
  uint64_t* dq_full = kv_empty + 1;        // dQ of the group's tiles complete (all key blocks)
  uint64_t* dq_empty = dq_full + 1;        // WG1 drained dQ (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L, nkv = p.nkv;
  // a CTA owns whole heads (g = blockIdx.x, + gridDim.x, ...) and walks each head's tile groups in order, so the
  // second group's read-add-write of dK / dV follows the first group's stores in the same thread
  const int heads_local = (p.heads_total - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                          static_cast<int>(gridDim.x);
  const int n_local = heads_local * p.ngroups;
  auto tiles_in_group = [&](int tg) { return min(2, p.nq - 2 * tg); };
  auto head_of = [&](int k) { return static_cast<int>(blockIdx.x) + (k / p.ngroups) * static_cast<int>(gridDim.x); };

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_o);
    for (int s = 0; s < BL_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(s_full, 1); mbar_init(pds_full, 128); mbar_init(mma2_done, 1);
    mbar_init(kv_full, 1); mbar_init(kv_empty, 128); mbar_init(dq_full, 1); mbar_init(dq_empty, 128);
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  // TMEM columns: S [0,128) dP [128,256) dV_j [256,320) dK_j [320,384) dQ_tt [384 + 64 tt, +64)

  if (warp == 8) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int bp = 0;
      for (int k = 0; k < n_local; ++k) {
        const int g = head_of(k), tg = k % p.ngroups;
        const int n = g / p.H, h = g - n * p.H;
        const int nt = tiles_in_group(tg);
        for (int j = 0; j < nkv; ++j) {
          for (int tt = 0; tt < nt; ++tt, ++bp) {
            const int t = 2 * tg + tt;
            const int s = bp % BL_STAGES;
            mbar_wait(&empty[s], ((bp / BL_STAGES) & 1) ^ 1);
            uint8_t* st = smem + s * BL_STAGE_BYTES;
            mbar_expect_tx(&full[s], BL_STAGE_BYTES);
            tma_load_2d(st, &tm_q, &full[s], 0, g * L + t * 128);
            tma_load_2d(st + TILE_BYTES, &tm_k, &full[s], 0, g * L + j * 128);
            tma_load_2d(st + 2 * TILE_BYTES, &tm_v, &full[s], 0, g * L + j * 128);
            tma_load_4d(st + 3 * TILE_BYTES, &tm_do, &full[s], 0, h, n, t * 128);
            tma_load_4d(st + 4 * TILE_BYTES, &tm_o, &full[s], 0, h, n, t * 128);
          }
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
    constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64) | IDESC_A_MN | IDESC_B_MN;  // A^T B forms
    constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64) | IDESC_B_MN;
    int n_pairs = 0;
    for (int k = 0; k < n_local; ++k) n_pairs += nkv * tiles_in_group(k % p.ngroups);
    auto issue_mma1 = [&](int bp) {
      const int s = bp % BL_STAGES;
      mbar_wait(&full[s], (bp / BL_STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = smem_u32(smem + s * BL_STAGE_BYTES);
        const uint64_t dq = umma_desc_kmajor_sw128(st), dk = umma_desc_kmajor_sw128(st + TILE_BYTES);
        const uint64_t dv = umma_desc_kmajor_sw128(st + 2 * TILE_BYTES), ddo = umma_desc_kmajor_sw128(st + 3 * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem_base, dq + 2 * kk, dk + 2 * kk, idesc_s, kk != 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem_base + 128, ddo + 2 * kk, dv + 2 * kk, idesc_s, kk != 0);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    if (n_pairs > 0) issue_mma1(0);
    int bp = 0, kvc = 0;  // pairs issued, key blocks completed
    for (int k = 0; k < n_local; ++k) {
      const int nt = tiles_in_group(k % p.ngroups);
      for (int j = 0; j < nkv; ++j) {
        for (int tt = 0; tt < nt; ++tt, ++bp) {
          const int s = bp % BL_STAGES;
          mbar_wait(pds_full, bp & 1);
          if (tt == 0) mbar_wait(kv_empty, (kvc & 1) ^ 1);            // dK / dV accumulators drained (previous block)
          if (j == 0 && tt == 0) mbar_wait(dq_empty, (k & 1) ^ 1);    // dQ accumulators drained (previous item)
          tc_fence_after();
          if (elect_one()) {
            const uint32_t st = smem_u32(smem + s * BL_STAGE_BYTES);
            const uint32_t aP = smem_u32(sP), aS = smem_u32(sdS);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)  // dV_j (+)= P^T dO_t   (K = query rows, 16 per step)
              umma_bf16_ss(tmem_base + 256, umma_desc_mnmajor_sw128(aP + kk * 2048, TILE_BYTES),
                           umma_desc_mnmajor_sw128(st + 3 * TILE_BYTES + kk * 2048, TILE_BYTES), idesc_t, (tt | kk) != 0);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)  // dK_j (+)= dS^T Q_t
              umma_bf16_ss(tmem_base + 320, umma_desc_mnmajor_sw128(aS + kk * 2048, TILE_BYTES),
                           umma_desc_mnmajor_sw128(st + kk * 2048, TILE_BYTES), idesc_t, (tt | kk) != 0);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)  // dQ_t (+)= dS K_j     (K = keys)
              umma_bf16_ss(tmem_base + 384 + tt * 64, umma_desc_kmajor_sw128(aS + (kk >> 2) * TILE_BYTES) + 2 * (kk & 3),
                           umma_desc_mnmajor_sw128(st + TILE_BYTES + kk * 2048, TILE_BYTES), idesc_q, (j | kk) != 0);
            umma_commit(mma2_done);
            umma_commit(&empty[s]);
            if (tt == nt - 1) umma_commit(kv_full);
            if (tt == nt - 1 && j == nkv - 1) umma_commit(dq_full);
          }
          __syncwarp();
          if (tt == nt - 1) ++kvc;
          if (bp + 1 < n_pairs) issue_mma1(bp + 1);
        }
      }
    }
  } else {
    const int wg = warp >> 2, quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    if (wg == 0) {
      // ---------------------------------------------------------- WG0: P, delta, dS
      int bp = 0;
      for (int k = 0; k < n_local; ++k) {
        const int g = head_of(k), tg = k % p.ngroups;
        const int nt = tiles_in_group(tg);
        for (int j = 0; j < nkv; ++j) {
          const int nvalid = min(128, L - j * 128);
          for (int tt = 0; tt < nt; ++tt, ++bp) {
            const int lq = (2 * tg + tt) * 128 + row;
            const bool valid = lq < L && g < p.heads_total;
            const float lse_s = valid ? p.lse[static_cast<size_t>(g) * L + lq] * LOG2E : 0.f;
            const int s = bp % BL_STAGES;
            mbar_wait(&full[s], (bp / BL_STAGES) & 1);  // dO_t, O_t tiles of this pair are in shared memory
            // delta = sum_d dO[row][d] * O[row][d]: both tiles share the swizzle, so matching physical chunks pair up
            float delta = 0.f;
            {
              const uint8_t* pdo = smem + s * BL_STAGE_BYTES + 3 * TILE_BYTES + row * 128;
              const uint8_t* po = pdo + TILE_BYTES;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint4 a = *reinterpret_cast<const uint4*>(pdo + c * 16);
                const uint4 b = *reinterpret_cast<const uint4*>(po + c * 16);
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float2 fa = unpack_bf16(aw[q]), fb = unpack_bf16(bw[q]);
                  delta = fmaf(fa.x, fb.x, fmaf(fa.y, fb.y, delta));
                }
              }
            }
            mbar_wait(s_full, bp & 1);
            tc_fence_after();
            if (bp > 0) mbar_wait(mma2_done, (bp - 1) & 1);  // MMA2 of the previous pair is done reading P / dS
#pragma unroll 1
            for (int c = 0; c < 128; c += 32) {
              uint32_t sv[32], dv[32];
              tmem_ld_32x32(t_lane + c, sv);
              tmem_ld_32x32(t_lane + 128 + c, dv);
              tmem_ld_wait();
              float pj[32], ds[32];
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) {
                const float e = (valid && c + jj < nvalid) ? fast_exp2(fmaf(__uint_as_float(sv[jj]), LOG2E, -lse_s)) : 0.f;
                pj[jj] = e;
                ds[jj] = e * (__uint_as_float(dv[jj]) - delta);
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int kc = (c >> 3) + q;
                const int half = kc >> 3, ch = kc & 7;
                const int off = half * TILE_BYTES + row * 128 + ((ch ^ (row & 7)) << 4);
                *reinterpret_cast<uint4*>(sP + off) =
                    make_uint4(pack_bf16(pj[8 * q], pj[8 * q + 1]), pack_bf16(pj[8 * q + 2], pj[8 * q + 3]),
                               pack_bf16(pj[8 * q + 4], pj[8 * q + 5]), pack_bf16(pj[8 * q + 6], pj[8 * q + 7]));
                *reinterpret_cast<uint4*>(sdS + off) =
                    make_uint4(pack_bf16(ds[8 * q], ds[8 * q + 1]), pack_bf16(ds[8 * q + 2], ds[8 * q + 3]),
                               pack_bf16(ds[8 * q + 4], ds[8 * q + 5]), pack_bf16(ds[8 * q + 6], ds[8 * q + 7]));
              }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(pds_full);
          }
        }
      }
    } else {
      // ---------------------------------------------------------- WG1: gradients out
      const size_t plane = static_cast<size_t>(p.heads_total) * L * 64;
      int kvc = 0;
      for (int k = 0; k < n_local; ++k) {
        const int g = head_of(k), tg = k % p.ngroups;
        const int nt = tiles_in_group(tg);
        const int n = g / p.H, h = g - n * p.H;
        const bool add_prev = tg > 0;  // a previous group of this head already wrote its dK / dV share
        for (int j = 0; j < nkv; ++j, ++kvc) {
          const int lk = j * 128 + row;
          const bool valid = lk < L && g < p.heads_total;
          bf16* tok = p.dqkv + (static_cast<size_t>(lk) * p.NB + n) * p.ld + h * 64;
          bf16* hm = p.ddelta != nullptr ? p.ddelta + plane + (static_cast<size_t>(g) * L + lk) * 64 : nullptr;  // dV' plane
          mbar_wait(kv_full, kvc & 1);
          tc_fence_after();
#pragma unroll 1
          for (int part = 0; part < 2; ++part) {  // 0: dV', 1: dK
            bf16* dst_tok = tok + (part == 0 ? 2 * p.D : p.D);
#pragma unroll
            for (int c = 0; c < 64; c += 32) {
              uint32_t v[32];
              tmem_ld_32x32(t_lane + 256 + part * 64 + c, v);
              tmem_ld_wait();
              if (valid) {
#pragma unroll
                for (int jj = 0; jj < 32; jj += 8) {
                  float f[8];
#pragma unroll
                  for (int t8 = 0; t8 < 8; ++t8) f[t8] = __uint_as_float(v[jj + t8]);
                  if (add_prev) add_bf16x8(f, *reinterpret_cast<const uint4*>(dst_tok + c + jj));
                  const uint4 o = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                             pack_bf16(f[6], f[7]));
                  *reinterpret_cast<uint4*>(dst_tok + c + jj) = o;
                  if (part == 0 && hm != nullptr) *reinterpret_cast<uint4*>(hm + c + jj) = o;
                }
              }
            }
          }
          tc_fence_before();
          mbar_arrive(kv_empty);
        }
        mbar_wait(dq_full, k & 1);
        tc_fence_after();
        for (int tt = 0; tt < nt; ++tt) {
          const int lq = (2 * tg + tt) * 128 + row;
          const bool valid = lq < L && g < p.heads_total;
          bf16* dst_tok = p.dqkv + (static_cast<size_t>(lq) * p.NB + n) * p.ld + h * 64;
          bf16* dst_hm = p.ddelta != nullptr ? p.ddelta + (static_cast<size_t>(g) * L + lq) * 64 : nullptr;
#pragma unroll
          for (int c = 0; c < 64; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32(t_lane + 384 + tt * 64 + c, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
              for (int jj = 0; jj < 32; jj += 8) {
                float f[8];
#pragma unroll
                for (int t8 = 0; t8 < 8; ++t8) f[t8] = __uint_as_float(v[jj + t8]);
                *reinterpret_cast<uint4*>(dst_tok + c + jj) =
                    make_uint4(pack_bf16(f[0] * 0.125f, f[1] * 0.125f), pack_bf16(f[2] * 0.125f, f[3] * 0.125f),
                               pack_bf16(f[4] * 0.125f, f[5] * 0.125f), pack_bf16(f[6] * 0.125f, f[7] * 0.125f));
                if (dst_hm != nullptr)
                  *reinterpret_cast<uint4*>(dst_hm + c + jj) = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]),
                                                                         pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(dq_empty);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool attn_tc_long_supported(const AttnShape& a) { return a.r == 0 && a.L > 128 && a.L <= 128 * MAX_BLOCKS && a.H * 64 == a.D; }

int attn_fwd_tc_long(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, bf16* o_tok,
                     float* lse) {
  PEVIT_REQUIRE(attn_tc_long_supported(a), "attn_fwd_tc_long: unsupported shape L=%d D=%d H=%d r=%d", a.L, a.D, a.H, a.r);
  const int heads = a.NB * a.H;
  const int nb = (a.L + 127) / 128;
  CUtensorMap tq, tk, tv;
  const uint64_t rows = static_cast<uint64_t>(heads) * a.L;
  if (make_tmap_bf16_2d(&tq, q, rows, 64, 64, 128, 64) != 0) return -1;
  if (make_tmap_bf16_2d(&tk, k, rows, 64, 64, 128, 64) != 0) return -1;
  if (make_tmap_bf16_2d(&tv, v, rows, 64, 64, 128, 64) != 0) return -1;
  FwdLongParams p{a.L, a.NB, a.H, a.D, heads, nb, nb, heads * nb, o_tok, lse};
  const int grid = p.num_items < sm_count() ? p.num_items : sm_count();
  static bool configured[64] = {};
  int dev = 0;
  PEVIT_CHECK_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    PEVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FL_SMEM));
    configured[dev & 63] = true;
  }
  ProfScope prof(s, PC_ATTN_FWD);
  PEVIT_CHECK_CUDA(launch_kernel(attn_fwd_long_kernel, dim3(grid), dim3(FL_THREADS), FL_SMEM, s, 1, tq, tk, tv, p));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int attn_bwd_tc_long(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, const bf16* o_tok,
                     const bf16* do_tok, const float* lse, bf16* dqkv, int ld_dqkv, bf16* ddelta) {
  PEVIT_REQUIRE(attn_tc_long_supported(a), "attn_bwd_tc_long: unsupported shape L=%d D=%d H=%d r=%d", a.L, a.D, a.H, a.r);
  PEVIT_REQUIRE(ld_dqkv % 8 == 0, "attn_bwd_tc_long: ld_dqkv=%d must be a multiple of 8", ld_dqkv);
  const int heads = a.NB * a.H;
  const int nb = (a.L + 127) / 128;
  const int ngroups = (nb + 1) / 2;
  CUtensorMap tq, tk, tv, tdo, to;
  const uint64_t rows = static_cast<uint64_t>(heads) * a.L;
  if (make_tmap_bf16_2d(&tq, q, rows, 64, 64, 128, 64) != 0) return -1;
  if (make_tmap_bf16_2d(&tk, k, rows, 64, 64, 128, 64) != 0) return -1;
  if (make_tmap_bf16_2d(&tv, v, rows, 64, 64, 128, 64) != 0) return -1;
  if (make_tmap_bf16_tok_heads(&tdo, do_tok, a.L, a.NB, a.H, a.D, 128) != 0) return -1;
  if (make_tmap_bf16_tok_heads(&to, o_tok, a.L, a.NB, a.H, a.D, 128) != 0) return -1;
  BwdLongParams p{a.L, a.NB, a.H, a.D, heads, nb, nb, ngroups, heads * ngroups, ld_dqkv, lse, dqkv, ddelta};
  const int grid = heads < sm_count() ? heads : sm_count();  // CTAs own whole heads
  static bool configured[64] = {};
  int dev = 0;
  PEVIT_CHECK_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    PEVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BL_SMEM));
    configured[dev & 63] = true;
  }
  ProfScope prof(s, PC_ATTN_BWD);
  PEVIT_CHECK_CUDA(launch_kernel(attn_bwd_long_kernel, dim3(grid), dim3(BL_THREADS), BL_SMEM, s, 1, tq, tk, tv, tdo, to, p));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

}  // namespace pevit
