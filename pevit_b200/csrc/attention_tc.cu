// Attention core on tcgen05 / TMEM (reference evaluation/model.py:803-815: bmm(q,k^T), softmax,
// bmm(p,v), head merge) for the head-major bf16 q', k, v' the in-projection GEMM (+ delta GEMM)
// leaves in HBM.  HBM-bound (arithmetic intensity L/2 FLOP/B), so the design goal is to keep
// ~all SMs streaming tiles: TMA-fed multi-stage ring, S and O accumulators double-buffered in
// TMEM, softmax by two warpgroups that alternate tiles (one thread per query row: row max / sum
// need no shuffles), P staged as the bf16 A operand of the second MMA, scores never touch HBM.
//
// Tile = 128 query rows = PACK heads (PACK = 2 for L <= 64: ViT-B/32, L = 50; PACK = 1 for
// 64 < L <= 128).  With PACK = 2 the two heads sit at rows 0.. and 64.. of every operand tile;
// S = Q K^T is then block-diagonal and each softmax thread only reads its own 64-column block,
// P's off-diagonal blocks stay zero so O = P V is exact.
#include "common.cuh"
#include "kernels.h"

namespace pevit {
namespace {

constexpr int TC_STAGES = 3;
constexpr int TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16
constexpr int STAGE_BYTES = 3 * TILE_BYTES;
constexpr int P_BYTES = 2 * TILE_BYTES;  // 128 rows x 128 keys
constexpr int FWD_SMEM = TC_STAGES * STAGE_BYTES + 2 * P_BYTES + 256 + 1024;
constexpr int FWD_THREADS = 320;  // 8 softmax warps, 1 TMA warp, 1 MMA warp
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct FwdParams {
  int L, NB, H, D, heads_total, num_tiles;
  bf16* o_tok;
  float* lse;
};

template <int PACK>
__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sP = smem + TC_STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * P_BYTES);
  uint64_t* full = bars;                   // [TC_STAGES]
  uint64_t* empty = full + TC_STAGES;      // [TC_STAGES]
  uint64_t* s_full = empty + TC_STAGES;    // [2]
  uint64_t* p_full = s_full + 2;           // [2]
  uint64_t* o_full = p_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L;
  const int n_local = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                      static_cast<int>(gridDim.x);

  // zero every operand / P buffer once: TMA only ever writes rows < L of each head slot, so the pad
  // rows (and P's off-diagonal blocks) stay exactly zero for the lifetime of the CTA.
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = (TC_STAGES * STAGE_BYTES + 2 * P_BYTES) / 16;
    for (int i = threadIdx.x; i < n16; i += FWD_THREADS) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&p_full[b], 128); mbar_init(&o_full[b], 1); }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();  // zero fill (generic proxy) before TMA / UMMA (async proxy) touch the buffers
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  // TMEM columns: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384)

  if (warp == 8) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int it = 0; it < n_local; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int s = it % TC_STAGES;
        const uint32_t ph = (it / TC_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        const int g0 = tile * PACK;
        const int nheads = min(PACK, p.heads_total - g0);
        uint8_t* st = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full[s], static_cast<uint32_t>(nheads) * 3u * static_cast<uint32_t>(L) * 128u);
        for (int j = 0; j < nheads; ++j) {
          const int row = (g0 + j) * L;
          tma_load_2d(st + j * 8192, &tm_q, &full[s], 0, row);
          tma_load_2d(st + TILE_BYTES + j * 8192, &tm_k, &full[s], 0, row);
          tma_load_2d(st + 2 * TILE_BYTES + j * 8192, &tm_v, &full[s], 0, row);
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64) | IDESC_B_MN;
    auto issue_s = [&](int it) {
      const int s = it % TC_STAGES, b = it & 1;
      mbar_wait(&full[s], (it / TC_STAGES) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sq = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t dq = umma_desc_kmajor_sw128(sq);
        const uint64_t dk = umma_desc_kmajor_sw128(sq + TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + b * 128, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[b]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int it) {
      const int s = it % TC_STAGES, b = it & 1;
      mbar_wait(&p_full[b], (it >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sp = smem_u32(sP + b * P_BYTES);
        const uint32_t sv = smem_u32(smem + s * STAGE_BYTES + 2 * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // A = P: K-major, keys 0..63 in the first [128][64] half-tile, 64..127 in the second
          const uint64_t da = umma_desc_kmajor_sw128(sp + (k >> 2) * TILE_BYTES) + 2 * (k & 3);
          // B = V: MN-major (64 d contiguous per key row); 16 keys per step = 2048 B
          const uint64_t db = umma_desc_mnmajor_sw128(sv + k * 2048, 8192);
          umma_bf16_ss(tmem_base + 256 + b * 64, da, db, idesc_o, k != 0);
        }
        umma_commit(&o_full[b]);
        umma_commit(&empty[s]);
      }
      __syncwarp();
    };
    if (n_local > 0) issue_s(0);
    for (int it = 0; it < n_local; ++it) {
      if (it + 1 < n_local) issue_s(it + 1);
      issue_pv(it);
    }
  } else {
    // ------------------------------------------------------------ softmax / epilogue warpgroups
    const int grp = warp >> 2;              // 0 or 1: handles local items it = grp, grp+2, ...
    const int quad = warp & 3;
    const int row = quad * 32 + lane;       // accumulator row == TMEM lane
    const int slot = PACK == 2 ? (row >> 6) : 0;   // which head of the pack this row belongs to
    const int l = PACK == 2 ? (row & 63) : row;    // token index within the head
    const int col0 = PACK == 2 ? slot * 64 : 0;    // first key column of this row's block
    constexpr int NCOL = PACK == 2 ? 64 : 128;
    uint8_t* myP = sP + grp * P_BYTES;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    for (int it = grp; it < n_local; it += 2) {
      const int tile = blockIdx.x + it * gridDim.x;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&s_full[grp], ph);
      tc_fence_after();
      // ---- scores of this row: NCOL fp32 values
      float sc[NCOL];
#pragma unroll
      for (int c = 0; c < NCOL; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_lane + grp * 128 + col0 + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) sc[c + j] = __uint_as_float(v[j]);
      }
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < NCOL; ++j) if (j < L) mx = fmaxf(mx, sc[j]);
      float sum = 0.f;
      const float mxs = mx * LOG2E;
#pragma unroll
      for (int j = 0; j < NCOL; ++j) {
        const float e = (j < L) ? fast_exp2(fmaf(sc[j], LOG2E, -mxs)) : 0.f;
        sc[j] = e;
        sum += e;
      }
      // ---- P (unnormalised, bf16) -> smem as the K-major A operand of O = P V
#pragma unroll
      for (int c = 0; c < NCOL / 8; ++c) {
        const uint4 pk = make_uint4(pack_bf16(sc[8 * c], sc[8 * c + 1]), pack_bf16(sc[8 * c + 2], sc[8 * c + 3]),
                                    pack_bf16(sc[8 * c + 4], sc[8 * c + 5]), pack_bf16(sc[8 * c + 6], sc[8 * c + 7]));
        const int kc = (col0 >> 3) + c;            // 16-byte chunk index along the 128 keys
        const int half = kc >> 3, ch = kc & 7;     // which [128][64] half-tile, chunk within its 128-B row
        *reinterpret_cast<uint4*>(myP + half * TILE_BYTES + row * 128 + ((ch ^ (row & 7)) << 4)) = pk;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[grp]);
      // ---- O = P V, normalise, merge heads: bf16 rows of o_tok[(l*NB + n)][h*64 ..]
      mbar_wait(&o_full[grp], ph);
      tc_fence_after();
      const int g = tile * PACK + slot;
      const bool valid = (l < L) && (g < p.heads_total);
      const float inv = 1.f / sum;
      const int n = g / p.H, h = g - n * p.H;
      bf16* orow = p.o_tok + (static_cast<size_t>(l) * p.NB + n) * p.D + h * 64;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_lane + 256 + grp * 64 + c, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float f[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) f[t] = __uint_as_float(v[j + t]) * inv;
            *reinterpret_cast<uint4*>(orow + c + j) = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]),
                                                                  pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
          }
        }
      }
      if (valid) p.lse[static_cast<size_t>(g) * L + l] = mx + __logf(sum);
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================
// Backward.  Per 128-row tile (PACK heads):
//   MMA1:  S = Q' K^T,  dP = dO V'^T                                  (TMEM cols [0,128) and [128,256))
//   WG0 :  P = exp(S - lse), delta = sum_j P dP, dS = P (dP - delta)  -> bf16 P, dS tiles in smem
//   MMA2:  dV' = P^T dO, dK = dS^T Q', dQ' = dS K                     (TMEM cols [256,320) [320,384) [384,448))
//   WG1 :  dQ'/8, dK, dV' -> token-major dqkv rows; dQ', dV' -> head-major d(delta) (F4: same memory)
// P^T and dS^T are not materialised: the [row][key] tiles are read as MN-major A operands, and dO, Q', K
// (64 contiguous d per row) as MN-major B operands.  delta uses the same P and dP that build dS, so the
// bf16 rounding of O never enters (and O is not read at all).
constexpr int BWD_STAGES = 2;
constexpr int BWD_STAGE_BYTES = 4 * TILE_BYTES;  // Q', K, V', dO
constexpr int BWD_SMEM = BWD_STAGES * BWD_STAGE_BYTES + 2 * P_BYTES + 256 + 1024;
constexpr int BWD_THREADS = 320;  // WG0 (4 warps) + WG1 (4 warps) + TMA warp + MMA warp

struct BwdParams {
  int L, NB, H, D, heads_total, num_tiles, ld;
  const float* lse;
  bf16* dqkv;
  bf16* ddelta;  // nullable
};

template <int PACK>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do, BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sP = smem + BWD_STAGES * BWD_STAGE_BYTES;
  uint8_t* sdS = sP + P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + P_BYTES);
  uint64_t* full = bars;                 // [BWD_STAGES]
  uint64_t* empty = full + BWD_STAGES;   // [BWD_STAGES]
  uint64_t* s_full = empty + BWD_STAGES; // MMA1 done
  uint64_t* pds_full = s_full + 1;       // WG0 wrote P, dS
  uint64_t* o2_full = pds_full + 1;      // MMA2 done
  uint64_t* o2_empty = o2_full + 1;      // WG1 drained dQ/dK/dV
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o2_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L;
  const int n_local = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                      static_cast<int>(gridDim.x);
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = (BWD_STAGES * BWD_STAGE_BYTES + 2 * P_BYTES) / 16;
    for (int i = threadIdx.x; i < n16; i += BWD_THREADS) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_do);
    for (int s = 0; s < BWD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(s_full, 1); mbar_init(pds_full, 128); mbar_init(o2_full, 1); mbar_init(o2_empty, 128);
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 8) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int it = 0; it < n_local; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int s = it % BWD_STAGES;
        mbar_wait(&empty[s], ((it / BWD_STAGES) & 1) ^ 1);
        const int g0 = tile * PACK;
        const int nheads = min(PACK, p.heads_total - g0);
        uint8_t* st = smem + s * BWD_STAGE_BYTES;
        mbar_expect_tx(&full[s], static_cast<uint32_t>(nheads) * 4u * static_cast<uint32_t>(L) * 128u);
        for (int j = 0; j < nheads; ++j) {
          const int g = g0 + j, row = g * L;
          const int n = g / p.H, h = g - n * p.H;
          tma_load_2d(st + j * 8192, &tm_q, &full[s], 0, row);
          tma_load_2d(st + TILE_BYTES + j * 8192, &tm_k, &full[s], 0, row);
          tma_load_2d(st + 2 * TILE_BYTES + j * 8192, &tm_v, &full[s], 0, row);
          tma_load_4d(st + 3 * TILE_BYTES + j * 8192, &tm_do, &full[s], 0, h, n, 0);
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
    constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64) | IDESC_A_MN | IDESC_B_MN;  // A^T B forms
    constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64) | IDESC_B_MN;
    auto issue_mma1 = [&](int it) {
      const int s = it % BWD_STAGES;
      mbar_wait(&full[s], (it / BWD_STAGES) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t st = smem_u32(smem + s * BWD_STAGE_BYTES);
        const uint64_t dq = umma_desc_kmajor_sw128(st), dk = umma_desc_kmajor_sw128(st + TILE_BYTES);
        const uint64_t dv = umma_desc_kmajor_sw128(st + 2 * TILE_BYTES), ddo = umma_desc_kmajor_sw128(st + 3 * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + 128, ddo + 2 * k, dv + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    if (n_local > 0) issue_mma1(0);
    for (int it = 0; it < n_local; ++it) {
      const int s = it % BWD_STAGES;
      mbar_wait(pds_full, it & 1);
      mbar_wait(o2_empty, (it & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t st = smem_u32(smem + s * BWD_STAGE_BYTES);
        const uint32_t aP = smem_u32(sP), aS = smem_u32(sdS);
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // dV' = P^T dO   (K = query rows, 16 per step)
          umma_bf16_ss(tmem_base + 256, umma_desc_mnmajor_sw128(aP + k * 2048, TILE_BYTES),
                       umma_desc_mnmajor_sw128(st + 3 * TILE_BYTES + k * 2048, TILE_BYTES), idesc_t, k != 0);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // dK = dS^T Q'
          umma_bf16_ss(tmem_base + 320, umma_desc_mnmajor_sw128(aS + k * 2048, TILE_BYTES),
                       umma_desc_mnmajor_sw128(st + k * 2048, TILE_BYTES), idesc_t, k != 0);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // dQ' = dS K     (K = keys)
          umma_bf16_ss(tmem_base + 384, umma_desc_kmajor_sw128(aS + (k >> 2) * TILE_BYTES) + 2 * (k & 3),
                       umma_desc_mnmajor_sw128(st + TILE_BYTES + k * 2048, TILE_BYTES), idesc_q, k != 0);
        }
        umma_commit(o2_full);
        umma_commit(&empty[s]);
      }
      __syncwarp();
      if (it + 1 < n_local) issue_mma1(it + 1);
    }
  } else {
    const int wg = warp >> 2, quad = warp & 3;
    const int row = quad * 32 + lane;
    const int slot = PACK == 2 ? (row >> 6) : 0;
    const int l = PACK == 2 ? (row & 63) : row;
    const int col0 = PACK == 2 ? slot * 64 : 0;
    constexpr int NCOL = PACK == 2 ? 64 : 128;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    if (wg == 0) {
      // ---------------------------------------------------------- WG0: P, delta, dS
      for (int it = 0; it < n_local; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int g = tile * PACK + slot;
        const bool valid = (l < L) && (g < p.heads_total);
        const float lse_s = valid ? p.lse[static_cast<size_t>(g) * L + l] * LOG2E : 0.f;
        mbar_wait(s_full, it & 1);
        tc_fence_after();
        float delta = 0.f;
#pragma unroll 1
        for (int c = 0; c < NCOL; c += 32) {
          uint32_t sv[32], dv[32];
          tmem_ld_32x32(t_lane + col0 + c, sv);
          tmem_ld_32x32(t_lane + 128 + col0 + c, dv);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float pj = (valid && c + j < L) ? fast_exp2(fmaf(__uint_as_float(sv[j]), LOG2E, -lse_s)) : 0.f;
            delta = fmaf(pj, __uint_as_float(dv[j]), delta);
          }
        }
        if (it > 0) {  // MMA2 of the previous tile must be done reading the P / dS tiles
          mbar_wait(o2_full, (it - 1) & 1);
        }
#pragma unroll 1
        for (int c = 0; c < NCOL; c += 32) {
          uint32_t sv[32], dv[32];
          tmem_ld_32x32(t_lane + col0 + c, sv);
          tmem_ld_32x32(t_lane + 128 + col0 + c, dv);
          tmem_ld_wait();
          float pj[32], ds[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float e = (valid && c + j < L) ? fast_exp2(fmaf(__uint_as_float(sv[j]), LOG2E, -lse_s)) : 0.f;
            pj[j] = e;
            ds[j] = e * (__uint_as_float(dv[j]) - delta);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int kc = ((col0 + c) >> 3) + q;
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
    } else {
      // ---------------------------------------------------------- WG1: gradients out
      const size_t plane = static_cast<size_t>(p.heads_total) * L * 64;
      for (int it = 0; it < n_local; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int g = tile * PACK + slot;
        const bool valid = (l < L) && (g < p.heads_total);
        const int n = g / p.H, h = g - n * p.H;
        bf16* tok = p.dqkv + (static_cast<size_t>(l) * p.NB + n) * p.ld + h * 64;
        bf16* hm = p.ddelta != nullptr ? p.ddelta + (static_cast<size_t>(g) * L + l) * 64 : nullptr;
        mbar_wait(o2_full, it & 1);
        tc_fence_after();
#pragma unroll 1
        for (int part = 0; part < 3; ++part) {   // 0: dV', 1: dK, 2: dQ'
          const float sc = part == 2 ? 0.125f : 1.f;
          bf16* dst_tok = tok + (part == 0 ? 2 * p.D : (part == 1 ? p.D : 0));
          bf16* dst_hm = (hm != nullptr && part != 1) ? hm + (part == 0 ? plane : 0) : nullptr;
#pragma unroll
          for (int c = 0; c < 64; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32(t_lane + 256 + part * 64 + c, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                float f[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) f[t] = __uint_as_float(v[j + t]);
                *reinterpret_cast<uint4*>(dst_tok + c + j) =
                    make_uint4(pack_bf16(f[0] * sc, f[1] * sc), pack_bf16(f[2] * sc, f[3] * sc),
                               pack_bf16(f[4] * sc, f[5] * sc), pack_bf16(f[6] * sc, f[7] * sc));
                if (dst_hm != nullptr)
                  *reinterpret_cast<uint4*>(dst_hm + c + j) = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]),
                                                                         pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(o2_empty);
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

bool attn_tc_supported(const AttnShape& a) { return a.r == 0 && a.L >= 1 && a.L <= 128 && a.H * 64 == a.D; }

int attn_fwd_tc(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, bf16* o_tok,
                float* lse) {
  PEVIT_REQUIRE(attn_tc_supported(a), "attn_fwd_tc: unsupported shape L=%d D=%d H=%d r=%d", a.L, a.D, a.H, a.r);
  const int heads = a.NB * a.H;
  const int pack = a.L <= 64 ? 2 : 1;
  const int tiles = (heads + pack - 1) / pack;
  CUtensorMap tq, tk, tv;
  const uint64_t rows = static_cast<uint64_t>(heads) * a.L;
  if (make_tmap_bf16_2d(&tq, q, rows, 64, 64, a.L, 64) != 0) return -1;
  if (make_tmap_bf16_2d(&tk, k, rows, 64, 64, a.L, 64) != 0) return -1;
  if (make_tmap_bf16_2d(&tv, v, rows, 64, 64, a.L, 64) != 0) return -1;
  FwdParams p{a.L, a.NB, a.H, a.D, heads, tiles, o_tok, lse};
  const int grid = tiles < sm_count() ? tiles : sm_count();
  ProfScope prof(s, PC_ATTN_FWD);
  if (pack == 2) {
    PEVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    PEVIT_CHECK_CUDA(launch_kernel(attn_fwd_tc_kernel<2>, dim3(grid), dim3(FWD_THREADS), FWD_SMEM, s, 1, tq, tk, tv, p));
  } else {
    PEVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    PEVIT_CHECK_CUDA(launch_kernel(attn_fwd_tc_kernel<1>, dim3(grid), dim3(FWD_THREADS), FWD_SMEM, s, 1, tq, tk, tv, p));
  }
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int attn_bwd_tc(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, const bf16* do_tok,
                const float* lse, bf16* dqkv, int ld_dqkv, bf16* ddelta) {
  PEVIT_REQUIRE(attn_tc_supported(a), "attn_bwd_tc: unsupported shape L=%d D=%d H=%d r=%d", a.L, a.D, a.H, a.r);
  PEVIT_REQUIRE(ld_dqkv % 8 == 0, "attn_bwd_tc: ld_dqkv=%d must be a multiple of 8", ld_dqkv);
  const int heads = a.NB * a.H;
  const int pack = a.L <= 64 ? 2 : 1;
  const int tiles = (heads + pack - 1) / pack;
  CUtensorMap tq, tk, tv, tdo;
  const uint64_t rows = static_cast<uint64_t>(heads) * a.L;
  if (make_tmap_bf16_2d(&tq, q, rows, 64, 64, a.L, 64) != 0) return -1;
  if (make_tmap_bf16_2d(&tk, k, rows, 64, 64, a.L, 64) != 0) return -1;
  if (make_tmap_bf16_2d(&tv, v, rows, 64, 64, a.L, 64) != 0) return -1;
  if (make_tmap_bf16_tok_heads(&tdo, do_tok, a.L, a.NB, a.H, a.D, a.L) != 0) return -1;
  BwdParams p{a.L, a.NB, a.H, a.D, heads, tiles, ld_dqkv, lse, dqkv, ddelta};
  const int grid = tiles < sm_count() ? tiles : sm_count();
  ProfScope prof(s, PC_ATTN_BWD);
  if (pack == 2) {
    PEVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    PEVIT_CHECK_CUDA(launch_kernel(attn_bwd_tc_kernel<2>, dim3(grid), dim3(BWD_THREADS), BWD_SMEM, s, 1, tq, tk, tv, tdo, p));
  } else {
    PEVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    PEVIT_CHECK_CUDA(launch_kernel(attn_bwd_tc_kernel<1>, dim3(grid), dim3(BWD_THREADS), BWD_SMEM, s, 1, tq, tk, tv, tdo, p));
  }
  PEVIT_CHECK_LAUNCH();
  return 0;
}

}  // namespace pevit
