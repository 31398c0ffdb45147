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
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace pevit {
namespace {

// Which single-lane guard each role of the backward kernel uses (overridable for A/B builds).  History: while the P / dS
// staging stores were still generic ST (see smem_align1024), back-to-back MMA issue through elect.sync measured SLOWER
// than `lane == 0` here (61.4 vs 55.3 us at L = 50, N = 256); with shared-space stores both measure 53.2 us.
#ifndef BWD_TMA_ONE
#define BWD_TMA_ONE elect_one()
#endif
#ifndef BWD_MMA_ONE
#define BWD_MMA_ONE elect_one()
#endif
#ifndef BWD_WG1_ONE
#define BWD_WG1_ONE elect_one()
#endif

constexpr int TC_STAGES = 3;
constexpr int ATTN_PREFETCH = 5;  // tiles pulled into L2 beyond the shared-memory ring
constexpr int TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16
constexpr int STAGE_BYTES = 3 * TILE_BYTES;
constexpr int P_BYTES = 2 * TILE_BYTES;  // 128 rows x 128 keys
constexpr int FWD_STATS_BYTES = 4 * 128 * 8;  // (row max, row sum) of the last four tiles
constexpr int FWD_SMEM = TC_STAGES * STAGE_BYTES + 2 * P_BYTES + FWD_STATS_BYTES + 256 + 1024;
constexpr int FWD_THREADS = 320;  // softmax warpgroup, epilogue warpgroup, TMA warp, MMA warp
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct FwdParams {
  int L, NB, H, D, heads_total, num_tiles;
  int debug;  // diagnostics (PEVIT_ATTN_DEBUG): 1 = no operand loads, 2 = no softmax math, 4 = no O staging / store
  bf16* o_tok;
  float* lse;
  int causal;  // additive -inf mask above the diagonal (CLIP text tower, model.py:1139-1145): keys j <= l only
};

// Roles: warps 0-3 softmax (every tile: S_b -> P_b + row statistics), warps 4-7 epilogue (every tile: O_b ->
// normalise -> global), warp 8 TMA, warp 9 MMA.  The two warpgroups are pipeline STAGES, not alternating owners of
// whole tiles: the softmax of tile i+1 never waits for the P V product of tile i, so the per-tile latency chain
// (TMA -> S -> softmax -> P V -> store) is overlapped three deep and the kernel runs at the rate of its slowest
// stage instead of the sum of all of them.
template <int PACK>
__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sP = smem + TC_STAGES * STAGE_BYTES;
  float2* stats = reinterpret_cast<float2*>(sP + 2 * P_BYTES);  // [tile & 3][row] = (max, sum)
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stats) + FWD_STATS_BYTES);
  uint64_t* full = bars;                   // [TC_STAGES]
  uint64_t* empty = full + TC_STAGES;      // [TC_STAGES]
  uint64_t* s_full = empty + TC_STAGES;    // [2]
  uint64_t* p_full = s_full + 2;           // [2]  4 warp arrivals
  uint64_t* o_full = p_full + 2;           // [2]
  uint64_t* o_empty = o_full + 2;          // [2]  4 warp arrivals
  uint64_t* p_free = o_empty + 2;          // [2]  the O tile staged in P_b has left through TMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L;
  const int n_local = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                      static_cast<int>(gridDim.x);

  // One tile's operands into ring slot it % TC_STAGES (TMA warp, one elected lane).
  auto load_tile = [&](int it) {
    const int tile = blockIdx.x + it * gridDim.x;
    const int s = it % TC_STAGES;
    const int g0 = tile * PACK;
    const int nheads = min(PACK, p.heads_total - g0);
    uint8_t* st = smem + s * STAGE_BYTES;
    if (p.debug & 1) { mbar_arrive(&full[s]); return; }
    mbar_expect_tx(&full[s], static_cast<uint32_t>(nheads) * 3u * static_cast<uint32_t>(L) * 128u);
    for (int j = 0; j < nheads; ++j) {
      const int row = (g0 + j) * L;
      tma_load_2d(st + j * 8192, &tm_q, &full[s], 0, row);
      tma_load_2d(st + TILE_BYTES + j * 8192, &tm_k, &full[s], 0, row);
      tma_load_2d(st + 2 * TILE_BYTES + j * 8192, &tm_v, &full[s], 0, row);
    }
  };
  const int n_early = min(TC_STAGES, n_local);  // tiles whose loads are issued from the prologue
  if (warp == 8) {
    if (elect_one()) {
      tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_o);
      for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&s_full[b], 1); mbar_init(&p_full[b], 4); mbar_init(&o_full[b], 1); mbar_init(&o_empty[b], 4);
        mbar_init(&p_free[b], 1);
      }
      fence_mbar_init();
      // The first ring-full of tiles is requested NOW, under the zero fill and the TMEM allocation below: TMA writes rows
      // < L of a head slot, the fill only rows >= L, so the two never touch the same bytes, and nobody waits on `full`
      // before the __syncthreads() that follows.  (This was ~1.5 us of exposed load latency per launch of an 11-tile kernel.)
      pdl_wait();
      for (int it = 0; it < n_early; ++it) load_tile(it);
    }
    __syncwarp();
  }
  // zero what TMA never writes and the MMAs still read: the pad rows [L, slot rows) of every operand slot, and the P
  // buffers (their off-diagonal blocks / pad rows must be exact zeros for the lifetime of the CTA)
  {
    constexpr int SLOT_ROWS = PACK == 2 ? 64 : 128;
    uint4* zp = reinterpret_cast<uint4*>(sP);
    for (int i = threadIdx.x; i < 2 * P_BYTES / 16; i += FWD_THREADS) zp[i] = make_uint4(0, 0, 0, 0);
    const int pad16 = (SLOT_ROWS - L) * 8;                 // 16-byte chunks of one slot's pad rows
    const int nslots = TC_STAGES * 3 * PACK;               // slots are contiguous: stage -> operand -> head slot
    if (pad16 > 0)
      for (int i = threadIdx.x; i < nslots * pad16; i += FWD_THREADS) {
        const int slot = i / pad16, c = i - slot * pad16;
        reinterpret_cast<uint4*>(smem + slot * (SLOT_ROWS * 128) + L * 128)[c] = make_uint4(0, 0, 0, 0);
      }
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
    if (elect_one()) {
      // The ring holds three tiles, but a tile's HBM latency is several tile-times: pull the operands of the tiles
      // further ahead into L2 now, so that the ring's own loads are L2 hits.
      auto prefetch_tile = [&](int it_) {
        if (it_ >= n_local) return;
        const int g0_ = (blockIdx.x + it_ * gridDim.x) * PACK;
        for (int j = 0; j < PACK && g0_ + j < p.heads_total; ++j) {
          tma_prefetch_l2_2d(&tm_q, 0, (g0_ + j) * L);
          tma_prefetch_l2_2d(&tm_k, 0, (g0_ + j) * L);
          tma_prefetch_l2_2d(&tm_v, 0, (g0_ + j) * L);
        }
      };
      for (int it_ = TC_STAGES; it_ < 2 * TC_STAGES + ATTN_PREFETCH; ++it_) prefetch_tile(it_);
      for (int it = n_early; it < n_local; ++it) {   // tiles [0, n_early) were requested in the prologue
        const int s = it % TC_STAGES;
        const uint32_t ph = (it / TC_STAGES) & 1;
        prefetch_tile(it + TC_STAGES + ATTN_PREFETCH);
        mbar_wait(&empty[s], ph ^ 1);
        load_tile(it);
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
      if (elect_one()) {
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
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&p_full[b], ph);
      mbar_wait(&o_empty[b], ph ^ 1);  // the epilogue warpgroup has read O_b of tile it-2
      tc_fence_after();
      if (elect_one()) {
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
    const int wgrole = warp >> 2;           // 0: softmax, 1: epilogue
    const int quad = warp & 3;
    const int row = quad * 32 + lane;       // accumulator row == TMEM lane
    const int slot = PACK == 2 ? (row >> 6) : 0;   // which head of the pack this row belongs to
    const int l = PACK == 2 ? (row & 63) : row;    // token index within the head
    const int col0 = PACK == 2 ? slot * 64 : 0;    // first key column of this row's block
    constexpr int NCOL = PACK == 2 ? 64 : 128;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int lim = p.causal ? min(L, l + 1) : L;  // keys [0, lim) of this row's block take part in the softmax
    if (wgrole == 0) {
      // ---------------------------------------------------------- softmax warpgroup
      for (int it = 0; it < n_local; ++it) {
        const int b = it & 1;
        uint8_t* myP = sP + b * P_BYTES;
        mbar_wait(&s_full[b], (it >> 1) & 1);
        tc_fence_after();
        // ---- scores of this row: NCOL fp32 values
        // One warp per scheduler runs this code, so instruction-level parallelism is all the latency hiding there
        // is: both TMEM loads are in flight before the wait, and the row max / row sum run as four independent
        // chains instead of one NCOL-long dependent one.
        float sc[NCOL];
        {
          uint32_t v[NCOL];
#pragma unroll
          for (int c = 0; c < NCOL; c += 32) tmem_ld_32x32(t_lane + b * 128 + col0 + c, *reinterpret_cast<uint32_t(*)[32]>(&v[c]));
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < NCOL; ++j) sc[j] = __uint_as_float(v[j]);
        }
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < NCOL; ++j) if (j < lim) m4[j & 3] = fmaxf(m4[j & 3], sc[j]);
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        const float mxs = mx * LOG2E;
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
          const float e = (j < lim) ? fast_exp2(fmaf(sc[j], LOG2E, -mxs)) : 0.f;
          sc[j] = e;
          s4[j & 3] += e;
        }
        const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        // ---- P (unnormalised, bf16) -> smem as the K-major A operand of O = P V
        mbar_wait(&p_free[b], ((it >> 1) & 1) ^ 1);  // the O tile of tile it-2, staged in this buffer, has left
#pragma unroll
        for (int c = 0; c < NCOL / 8; ++c) {
          const uint4 pk = make_uint4(pack_bf16(sc[8 * c], sc[8 * c + 1]), pack_bf16(sc[8 * c + 2], sc[8 * c + 3]),
                                      pack_bf16(sc[8 * c + 4], sc[8 * c + 5]), pack_bf16(sc[8 * c + 6], sc[8 * c + 7]));
          const int kc = (col0 >> 3) + c;            // 16-byte chunk index along the 128 keys
          const int half = kc >> 3, ch = kc & 7;     // which [128][64] half-tile, chunk within its 128-B row
          *reinterpret_cast<uint4*>(myP + half * TILE_BYTES + row * 128 + ((ch ^ (row & 7)) << 4)) = pk;
        }
        stats[(it & 3) * 128 + row] = make_float2(mx, sum);  // read by the epilogue warpgroup after o_full
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[b]);  // one arrival per warp: 128 arrivals on one mbarrier serialise (~600 cycles)
      }
    } else {
      // ---------------------------------------------------------- epilogue warpgroup
      // The normalised O rows are staged in the part of P_b this row's softmax thread rewrites every tile anyway
      // (P_b is dead once O_b is complete) and leave as one TMA store per head: [L tokens][64] -> o_tok rows
      // (l * NB + n), columns h * 64.., coalesced instead of 128 scattered 16-byte stores per warp.
      auto elected = [&]() { return warp == 4 && elect_one(); };  // one lane of the warpgroup's first warp, the same every time
      for (int it = 0; it < n_local; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int b = it & 1;
        if (elected() && it > 0) {  // previous tile's store has read its staging buffer: hand P_(b^1) back to the softmax
          tma_store_wait_read<0>();
          mbar_arrive(&p_free[b ^ 1]);
        }
        mbar_wait(&o_full[b], (it >> 1) & 1);
        tc_fence_after();
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(t_lane + 256 + b * 64, v0);
        tmem_ld_32x32(t_lane + 256 + b * 64 + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[b]);  // O_b is in registers: the P V product of tile it+2 may overwrite it
        const float2 ms = stats[(it & 3) * 128 + row];
        const int g = tile * PACK + slot;
        const bool valid = (l < L) && (g < p.heads_total);
        uint8_t* stg = sP + b * P_BYTES + (PACK == 2 ? slot * TILE_BYTES : 0) + row * 128;
        if (valid && !(p.debug & 4)) {
          const float inv = 1.f / ms.y;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            *reinterpret_cast<uint4*>(stg + (((j >> 3) ^ (row & 7)) << 4)) =
                make_uint4(pack_bf16(__uint_as_float(v0[j]) * inv, __uint_as_float(v0[j + 1]) * inv),
                           pack_bf16(__uint_as_float(v0[j + 2]) * inv, __uint_as_float(v0[j + 3]) * inv),
                           pack_bf16(__uint_as_float(v0[j + 4]) * inv, __uint_as_float(v0[j + 5]) * inv),
                           pack_bf16(__uint_as_float(v0[j + 6]) * inv, __uint_as_float(v0[j + 7]) * inv));
            *reinterpret_cast<uint4*>(stg + ((((j >> 3) + 4) ^ (row & 7)) << 4)) =
                make_uint4(pack_bf16(__uint_as_float(v1[j]) * inv, __uint_as_float(v1[j + 1]) * inv),
                           pack_bf16(__uint_as_float(v1[j + 2]) * inv, __uint_as_float(v1[j + 3]) * inv),
                           pack_bf16(__uint_as_float(v1[j + 4]) * inv, __uint_as_float(v1[j + 5]) * inv),
                           pack_bf16(__uint_as_float(v1[j + 6]) * inv, __uint_as_float(v1[j + 7]) * inv));
          }
          p.lse[static_cast<size_t>(g) * L + l] = ms.x + __logf(ms.y);
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (elected() && !(p.debug & 4)) {
#pragma unroll
          for (int j = 0; j < PACK; ++j) {
            const int gj = tile * PACK + j;
            if (gj < p.heads_total) {
              const int n = gj / p.H, h = gj - n * p.H;
              tma_store_4d(&tm_o, sP + b * P_BYTES + (PACK == 2 ? j * (TILE_BYTES + 8192) : 0), 0, h, n, 0);
            }
          }
          tma_store_commit();
        }
      }
      if (elected()) tma_store_wait_all<0>();
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
// Backward.  Per 128-row tile (PACK heads), TMEM column set b = tile & 1 (256 columns each):
//   MMA1:  S = Q' K^T  -> set+[0,128),  dP = dO V'^T -> set+[128,256)
//   WG0 :  P = exp(S - lse), delta = sum_j P dP, dS = P (dP - delta)  -> bf16 P, dS tiles in smem
//   MMA2:  dV' = P^T dO -> set+[0,64), dK = dS^T Q' -> set+[64,128), dQ' = dS K -> set+[128,192)
//          (the gradients overwrite the S / dP columns WG0 has just consumed, which is what lets the two sets
//          double-buffer inside the 512 TMEM columns: MMA1 of tile i+1 runs while tile i is still in flight)
//   WG1 :  dQ'/8, dK, dV' -> token-major dqkv rows, staged in smem and TMA-stored per head (coalesced);
//          dQ', dV' -> head-major d(delta) (F4: same memory), contiguous per head so stored directly
// P^T and dS^T are not materialised: the [row][key] tiles are read as MN-major A operands, and dO, Q', K
// (64 contiguous d per row) as MN-major B operands.  delta uses the same P and dP that build dS, so the
// bf16 rounding of O never enters (and O is not read at all).
constexpr int BWD_STAGES = 2;
constexpr int BWD_STAGE_BYTES = 4 * TILE_BYTES;  // Q', K, V', dO
constexpr int BWD_STAGING_BYTES = 2 * TILE_BYTES;  // two [128 rows][64] bf16 boxes for the token-major gradient stores
constexpr int BWD_SMEM = BWD_STAGES * BWD_STAGE_BYTES + 2 * P_BYTES + BWD_STAGING_BYTES + 256 + 1024;
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
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                   const __grid_constant__ CUtensorMap tm_dqkv, BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sP = smem + BWD_STAGES * BWD_STAGE_BYTES;
  uint8_t* sdS = sP + P_BYTES;
  uint8_t* stg = sdS + P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + BWD_STAGING_BYTES);
  uint64_t* full = bars;                  // [BWD_STAGES]
  uint64_t* empty = full + BWD_STAGES;    // [BWD_STAGES]
  uint64_t* s_full = empty + BWD_STAGES;  // [2]  MMA1 done (per TMEM set)
  uint64_t* pds_full = s_full + 2;        // WG0 wrote P, dS (4 warp arrivals)
  uint64_t* o2_full = pds_full + 1;       // [2]  MMA2 done (per TMEM set)
  uint64_t* o2_empty = o2_full + 2;       // [2]  WG1 read dQ/dK/dV out of the set (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o2_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L;
  const int n_local = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                      static_cast<int>(gridDim.x);
  // One tile's operands (Q', K, V', dO) into ring slot it % BWD_STAGES (TMA warp, one elected lane).
  auto load_tile = [&](int it) {
    const int tile = blockIdx.x + it * gridDim.x;
    const int s = it % BWD_STAGES;
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
  };
  const int n_early = min(BWD_STAGES, n_local);  // tiles whose loads are issued from the prologue (see the forward kernel)
  if (warp == 8) {
    if (BWD_TMA_ONE) {
      tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_do);
      tma_prefetch_desc(&tm_dqkv);
      for (int s = 0; s < BWD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&o2_full[b], 1); mbar_init(&o2_empty[b], 4); }
      mbar_init(pds_full, 4);
      fence_mbar_init();
      pdl_wait();
      for (int it = 0; it < n_early; ++it) load_tile(it);
    }
    __syncwarp();
  }
  // zero the pad rows [L, slot rows) of every operand slot and the P / dS tiles (TMA only writes rows < L)
  {
    constexpr int SLOT_ROWS = PACK == 2 ? 64 : 128;
    uint4* zp = reinterpret_cast<uint4*>(sP);
    for (int i = threadIdx.x; i < 2 * P_BYTES / 16; i += BWD_THREADS) zp[i] = make_uint4(0, 0, 0, 0);
    const int pad16 = (SLOT_ROWS - L) * 8;
    const int nslots = BWD_STAGES * 4 * PACK;
    if (pad16 > 0)
      for (int i = threadIdx.x; i < nslots * pad16; i += BWD_THREADS) {
        const int slot = i / pad16, c = i - slot * pad16;
        reinterpret_cast<uint4*>(smem + slot * (SLOT_ROWS * 128) + L * 128)[c] = make_uint4(0, 0, 0, 0);
      }
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
    if (BWD_TMA_ONE) {
      auto prefetch_tile = [&](int it_) {  // see the forward kernel: L2 prefetch beyond the two-stage ring
        if (it_ >= n_local) return;
        const int g0_ = (blockIdx.x + it_ * gridDim.x) * PACK;
        for (int j = 0; j < PACK && g0_ + j < p.heads_total; ++j) {
          const int g_ = g0_ + j, n_ = g_ / p.H, h_ = g_ - n_ * p.H;
          tma_prefetch_l2_2d(&tm_q, 0, g_ * L);
          tma_prefetch_l2_2d(&tm_k, 0, g_ * L);
          tma_prefetch_l2_2d(&tm_v, 0, g_ * L);
          tma_prefetch_l2_4d(&tm_do, 0, h_, n_, 0);
        }
      };
      for (int it_ = BWD_STAGES; it_ < 2 * BWD_STAGES + ATTN_PREFETCH; ++it_) prefetch_tile(it_);
      for (int it = n_early; it < n_local; ++it) {   // tiles [0, n_early) were requested in the prologue
        const int s = it % BWD_STAGES;
        prefetch_tile(it + BWD_STAGES + ATTN_PREFETCH);
        mbar_wait(&empty[s], ((it / BWD_STAGES) & 1) ^ 1);
        load_tile(it);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
    constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64) | IDESC_A_MN | IDESC_B_MN;  // A^T B forms
    constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64) | IDESC_B_MN;
    auto issue_mma1 = [&](int it) {
      const int s = it % BWD_STAGES;
      const uint32_t set = tmem_base + (it & 1) * 256;
      mbar_wait(&full[s], (it / BWD_STAGES) & 1);
      tc_fence_after();
      if (BWD_MMA_ONE) {
        const uint32_t st = smem_u32(smem + s * BWD_STAGE_BYTES);
        const uint64_t dq = umma_desc_kmajor_sw128(st), dk = umma_desc_kmajor_sw128(st + TILE_BYTES);
        const uint64_t dv = umma_desc_kmajor_sw128(st + 2 * TILE_BYTES), ddo = umma_desc_kmajor_sw128(st + 3 * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(set, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(set + 128, ddo + 2 * k, dv + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[it & 1]);
      }
      __syncwarp();
    };
    if (n_local > 0) issue_mma1(0);
    for (int it = 0; it < n_local; ++it) {
      const int s = it % BWD_STAGES;
      const uint32_t set = tmem_base + (it & 1) * 256;
      mbar_wait(pds_full, it & 1);  // P, dS are in smem; S / dP of this set have been consumed
      tc_fence_after();
      if (BWD_MMA_ONE) {
        const uint32_t st = smem_u32(smem + s * BWD_STAGE_BYTES);
        const uint32_t aP = smem_u32(sP), aS = smem_u32(sdS);
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // dV' = P^T dO   (K = query rows, 16 per step)
          umma_bf16_ss(set, umma_desc_mnmajor_sw128(aP + k * 2048, TILE_BYTES),
                       umma_desc_mnmajor_sw128(st + 3 * TILE_BYTES + k * 2048, TILE_BYTES), idesc_t, k != 0);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // dK = dS^T Q'
          umma_bf16_ss(set + 64, umma_desc_mnmajor_sw128(aS + k * 2048, TILE_BYTES),
                       umma_desc_mnmajor_sw128(st + k * 2048, TILE_BYTES), idesc_t, k != 0);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // dQ' = dS K     (K = keys)
          umma_bf16_ss(set + 128, umma_desc_kmajor_sw128(aS + (k >> 2) * TILE_BYTES) + 2 * (k & 3),
                       umma_desc_mnmajor_sw128(st + TILE_BYTES + k * 2048, TILE_BYTES), idesc_q, k != 0);
        }
        umma_commit(&o2_full[it & 1]);
        umma_commit(&empty[s]);
      }
      __syncwarp();
      if (it + 1 < n_local) {
        // tile it+1 reuses the column set of tile it-1: its gradients must have been read out
        if (it >= 1) mbar_wait(&o2_empty[(it + 1) & 1], ((it - 1) >> 1) & 1);
        issue_mma1(it + 1);
      }
    }
  } else {
    const int wg = warp >> 2, quad = warp & 3;
    const int row = quad * 32 + lane;
    const int slot = PACK == 2 ? (row >> 6) : 0;
    const int l = PACK == 2 ? (row & 63) : row;
    const int col0 = PACK == 2 ? slot * 64 : 0;
    constexpr int NCOL = PACK == 2 ? 64 : 128;
    if (wg == 0) {
      // ---------------------------------------------------------- WG0: P, delta, dS
      auto load_lse = [&](int it_) -> float {
        const int g_ = (blockIdx.x + it_ * gridDim.x) * PACK + slot;
        return (it_ < n_local && l < L && g_ < p.heads_total) ? p.lse[static_cast<size_t>(g_) * L + l] : 0.f;
      };
      float lse_next = load_lse(0);  // fetched one tile ahead: its latency hides behind the previous tile's work
      for (int it = 0; it < n_local; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int g = tile * PACK + slot;
        const bool valid = (l < L) && (g < p.heads_total);
        const float lse_s = lse_next * LOG2E;
        lse_next = load_lse(it + 1);
        mbar_wait(&s_full[it & 1], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (it & 1) * 256;
        if constexpr (PACK == 2) {
          // one pass: the row's 64 scores and 64 dP values stay in registers between delta and dS
          float pr[64], dp[64];
          {
            uint32_t sv[64], dv[64];
#pragma unroll
            for (int c = 0; c < 64; c += 32) {
              tmem_ld_32x32(t_lane + col0 + c, *reinterpret_cast<uint32_t(*)[32]>(&sv[c]));
              tmem_ld_32x32(t_lane + 128 + col0 + c, *reinterpret_cast<uint32_t(*)[32]>(&dv[c]));
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 64; ++j) {
              pr[j] = (valid && j < L) ? fast_exp2(fmaf(__uint_as_float(sv[j]), LOG2E, -lse_s)) : 0.f;
              dp[j] = __uint_as_float(dv[j]);
            }
          }
          float d4[4] = {0.f, 0.f, 0.f, 0.f};  // four independent chains (see the forward softmax)
#pragma unroll
          for (int j = 0; j < 64; ++j) d4[j & 3] = fmaf(pr[j], dp[j], d4[j & 3]);
          const float delta = (d4[0] + d4[1]) + (d4[2] + d4[3]);
          if (it > 0) mbar_wait(&o2_full[(it - 1) & 1], ((it - 1) >> 1) & 1);  // MMA2 of the previous tile is done reading P / dS
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int kc = (col0 >> 3) + q;
            const int half = kc >> 3, ch = kc & 7;
            const int off = half * TILE_BYTES + row * 128 + ((ch ^ (row & 7)) << 4);
            float ds[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) ds[t] = pr[8 * q + t] * (dp[8 * q + t] - delta);
            *reinterpret_cast<uint4*>(sP + off) =
                make_uint4(pack_bf16(pr[8 * q], pr[8 * q + 1]), pack_bf16(pr[8 * q + 2], pr[8 * q + 3]),
                           pack_bf16(pr[8 * q + 4], pr[8 * q + 5]), pack_bf16(pr[8 * q + 6], pr[8 * q + 7]));
            *reinterpret_cast<uint4*>(sdS + off) = make_uint4(pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]),
                                                              pack_bf16(ds[4], ds[5]), pack_bf16(ds[6], ds[7]));
          }
        } else {
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
          if (it > 0) mbar_wait(&o2_full[(it - 1) & 1], ((it - 1) >> 1) & 1);  // MMA2 of the previous tile done reading P / dS
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
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pds_full);
      }
    } else {
      // ---------------------------------------------------------- WG1: gradients out
      const size_t plane = static_cast<size_t>(p.heads_total) * L * 64;
      auto elected = [&]() { return warp == 4 && BWD_WG1_ONE; };  // one lane of the warpgroup's first warp, the same every time
      int pc = 0;  // parts stored so far (staging box = pc & 1)
      for (int it = 0; it < n_local; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int g = tile * PACK + slot;
        const bool valid = (l < L) && (g < p.heads_total);
        bf16* hm = p.ddelta != nullptr ? p.ddelta + (static_cast<size_t>(g) * L + l) * 64 : nullptr;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (it & 1) * 256;
        mbar_wait(&o2_full[it & 1], (it >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int part = 0; part < 3; ++part, ++pc) {   // 0: dV' (columns 2D..), 1: dK (D..), 2: dQ' (0.., scaled 1/8)
          uint32_t v[64];
          tmem_ld_32x32(t_lane + part * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_ld_32x32(t_lane + part * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          tmem_ld_wait();
          if (part == 2) {  // the set's last columns are in registers: MMA1 of tile it+2 may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o2_empty[it & 1]);
          }
          // head-major d(delta) planes (dQ' -> plane 0, dV' -> plane 1): contiguous per head, stored directly
          if (valid && hm != nullptr && part != 1) {
            bf16* dst = hm + (part == 0 ? plane : 0);
#pragma unroll
            for (int j = 0; j < 64; j += 8)
              *reinterpret_cast<uint4*>(dst + j) =
                  make_uint4(pack_bf16(__uint_as_float(v[j]), __uint_as_float(v[j + 1])),
                             pack_bf16(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])),
                             pack_bf16(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])),
                             pack_bf16(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7])));
          }
          // token-major dqkv rows: staged [row][64] (swizzled) and TMA-stored per head
          uint8_t* box = stg + (pc & 1) * TILE_BYTES;
          if (elected()) tma_store_wait_read<1>();  // the store that last used this box (two parts ago) has drained
          named_bar_sync(2, 128);
          if (valid) {
            const float sc = part == 2 ? 0.125f : 1.f;
            uint8_t* rp = box + row * 128;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<uint4*>(rp + ((q ^ (row & 7)) << 4)) =
                  make_uint4(pack_bf16(__uint_as_float(v[8 * q]) * sc, __uint_as_float(v[8 * q + 1]) * sc),
                             pack_bf16(__uint_as_float(v[8 * q + 2]) * sc, __uint_as_float(v[8 * q + 3]) * sc),
                             pack_bf16(__uint_as_float(v[8 * q + 4]) * sc, __uint_as_float(v[8 * q + 5]) * sc),
                             pack_bf16(__uint_as_float(v[8 * q + 6]) * sc, __uint_as_float(v[8 * q + 7]) * sc));
          }
          fence_proxy_async_smem();
          named_bar_sync(2, 128);
          if (elected()) {
            const int hbase = part == 0 ? 2 * p.H : (part == 1 ? p.H : 0);  // dqkv columns: [dq | dk | dv] heads
#pragma unroll
            for (int j = 0; j < PACK; ++j) {
              const int gj = tile * PACK + j;
              if (gj < p.heads_total) {
                const int n = gj / p.H, h = gj - n * p.H;
                tma_store_4d(&tm_dqkv, box + j * 8192, 0, hbase + h, n, 0);
              }
            }
            tma_store_commit();
          }
        }
      }
      if (elected()) tma_store_wait_all<0>();
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
  CUtensorMap to;
  if (make_tmap_bf16_tok_heads(&to, o_tok, a.L, a.NB, a.H, a.D, a.L) != 0) return -1;
  static const int dbg = getenv("PEVIT_ATTN_DEBUG") ? atoi(getenv("PEVIT_ATTN_DEBUG")) : 0;
  FwdParams p{a.L, a.NB, a.H, a.D, heads, tiles, dbg, o_tok, lse, a.causal};
  const int grid = tiles < sm_count() ? tiles : sm_count();
  ProfScope prof(s, PC_ATTN_FWD);
  if (pack == 2) {
    PEVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    PEVIT_CHECK_CUDA(launch_kernel(attn_fwd_tc_kernel<2>, dim3(grid), dim3(FWD_THREADS), FWD_SMEM, s, 1, tq, tk, tv, to, p));
  } else {
    PEVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    PEVIT_CHECK_CUDA(launch_kernel(attn_fwd_tc_kernel<1>, dim3(grid), dim3(FWD_THREADS), FWD_SMEM, s, 1, tq, tk, tv, to, p));
  }
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int attn_bwd_tc(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, const bf16* do_tok,
                const float* lse, bf16* dqkv, int ld_dqkv, bf16* ddelta) {
  PEVIT_REQUIRE(attn_tc_supported(a), "attn_bwd_tc: unsupported shape L=%d D=%d H=%d r=%d", a.L, a.D, a.H, a.r);
  PEVIT_REQUIRE(ld_dqkv % 8 == 0, "attn_bwd_tc: ld_dqkv=%d must be a multiple of 8", ld_dqkv);
  PEVIT_REQUIRE(!a.causal, "attn_bwd_tc: the causal mask is forward-only (the text tower is frozen)");
  const int heads = a.NB * a.H;
  const int pack = a.L <= 64 ? 2 : 1;
  const int tiles = (heads + pack - 1) / pack;
  CUtensorMap tq, tk, tv, tdo;
  const uint64_t rows = static_cast<uint64_t>(heads) * a.L;
  if (make_tmap_bf16_2d(&tq, q, rows, 64, 64, a.L, 64) != 0) return -1;
  if (make_tmap_bf16_2d(&tk, k, rows, 64, 64, a.L, 64) != 0) return -1;
  if (make_tmap_bf16_2d(&tv, v, rows, 64, 64, a.L, 64) != 0) return -1;
  if (make_tmap_bf16_tok_heads(&tdo, do_tok, a.L, a.NB, a.H, a.D, a.L) != 0) return -1;
  CUtensorMap tdq;  // dqkv rows viewed as 3H heads of 64 columns (the low-rank columns beyond 3D are not covered)
  if (make_tmap_bf16_tok_heads(&tdq, dqkv, a.L, a.NB, 3 * a.H, ld_dqkv, a.L) != 0) return -1;
  BwdParams p{a.L, a.NB, a.H, a.D, heads, tiles, ld_dqkv, lse, dqkv, ddelta};
  const int grid = tiles < sm_count() ? tiles : sm_count();
  ProfScope prof(s, PC_ATTN_BWD);
  if (pack == 2) {
    PEVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    PEVIT_CHECK_CUDA(launch_kernel(attn_bwd_tc_kernel<2>, dim3(grid), dim3(BWD_THREADS), BWD_SMEM, s, 1, tq, tk, tv, tdo, tdq, p));
  } else {
    PEVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    PEVIT_CHECK_CUDA(launch_kernel(attn_bwd_tc_kernel<1>, dim3(grid), dim3(BWD_THREADS), BWD_SMEM, s, 1, tq, tk, tv, tdo, tdq, p));
  }
  PEVIT_CHECK_LAUNCH();
  return 0;
}

}  // namespace pevit
