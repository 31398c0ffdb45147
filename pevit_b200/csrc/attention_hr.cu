// "Head-resident" attention core on tcgen05 / TMEM for sequences longer than one 128-row tile (ViT-B/16: L = 197,
// ViT-L/14: L = 257; any 128 < L <= 384).  Reference: evaluation/model.py:803-815 (bmm(q, k^T), softmax, bmm(p, v),
// head merge) and its autograd; q', k, v' are head-major bf16 with the low-rank delta already applied.
//
// Work item = ONE (image, head): its Q, K, V (and dO, O in the backward) are brought into shared memory ONCE by
// TMA -- full 128-row tiles plus a tail tile of ceil16(L mod 128) rows, so HBM traffic is the algorithmic minimum and
// the padded part of the last tile costs no exponentials and almost no MMA columns -- and every (query tile, key
// block) pair of the head is computed from there.
//
//   forward   two stages of head operands (the next head loads while this one computes).  Per pair: S = Q_t K_j^T in
//             TMEM, one softmax thread per query row computes the block-local statistics (m_j, l_j) and writes the
//             bf16 probabilities BACK INTO TMEM over the S columns it has just read; O_j = P V_j is then a tcgen05.mma
//             with the A operand taken from TMEM (no shared-memory round trip, no proxy fence), one accumulator per
//             key block.  A third warpgroup merges the blocks exactly (O = sum_j e^{m_j-m} O_j / sum_j e^{m_j-m} l_j),
//             so the softmax warpgroups never wait for an epilogue.  Nothing is rescaled in TMEM, no partial touches HBM.
//   backward  see the second half of this file.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "trace.cuh"

namespace pevit {
namespace {

constexpr int TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16, 128B-swizzled
constexpr float LOG2E = 1.4426950408889634f;
constexpr int HR_MAXT = 3;             // tiles per head (L <= 384)

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Geometry of one head's operands in shared memory: nt tiles, the last one `tail` valid rows stored as tail16 rows.
struct HrGeom {
  int L, NB, H, D, heads, nt, tail, tail16;
  int tensor_bytes;  // (nt - 1) * TILE_BYTES + tail16 * 128
};
__host__ __device__ inline int hr_rows(const HrGeom& g, int t) { return t == g.nt - 1 ? g.tail : 128; }      // valid rows
__host__ __device__ inline int hr_rows16(const HrGeom& g, int t) { return t == g.nt - 1 ? g.tail16 : 128; }  // stored rows

HrGeom make_geom(const AttnShape& a) {
  HrGeom g{};
  g.L = a.L; g.NB = a.NB; g.H = a.H; g.D = a.D; g.heads = a.NB * a.H;
  g.nt = (a.L + 127) / 128;
  g.tail = a.L - 128 * (g.nt - 1);
  g.tail16 = (g.tail + 15) / 16 * 16;
  g.tensor_bytes = (g.nt - 1) * TILE_BYTES + g.tail16 * 128;
  return g;
}

// ===================================================================================================== forward
// Unit = (query tile t, key block jb).  The stored keys of a head (L16 = ceil16(L)) are cut into nkb = ceil(L16 / 112) blocks
// of <= 112 keys -- ViT-B/16: 112 + 96, ViT-L/14: 96 + 96 + 80 -- and each block is ONE MMA shape (N = block width), so
// the key tail costs neither a unit nor exponentials.
//   * Unit u belongs to softmax warpgroup u & 1, ALONE: one thread per query row reads the row's <= 112 scores from
//     TMEM once (a warp reads TMEM at ~48 B/clk however many loads are in flight, so a second pass would cost as much as
//     the exponentials), keeps them in registers (setmaxnreg moves registers from the other roles to the softmax
//     warpgroups), and writes the bf16 probabilities back over the scores.  No barrier couples the two warpgroups, so
//     their load / MUFU / store phases interleave on every SM sub-partition.
//   * Up to three score buffers rotate over the units (u % nbuf): the scores of unit u+2 are computed while the
//     warpgroups work on units u and u+1, which hides the MMA round trip.
//   * One accumulator O_jb per key block, merged exactly by a third warpgroup (O = sum_jb e^{m_jb - m} O_jb / l): nothing
//     is rescaled in TMEM and no partial result touches HBM.
constexpr int HF_THREADS = 512;  // softmax WG0, softmax WG1, merge WG, [TMA warp, MMA warp, two idle warps]
constexpr int HF_MAXKB = 4;      // key blocks per head
constexpr int HF_MAXCOLS = 112;  // keys per block = score columns per softmax thread
constexpr int HF_MAXBUF = 3;     // score buffers in TMEM
// Row statistics of the tiles in flight.  The softmax warpgroups run up to nbuf - 1 units ahead of the P V pointer, which
// itself waits for the merge warpgroup at every tile boundary: while the merge reads tile T-1 the softmax may already
// be writing tile T+1, so two slots are NOT enough (compute-sanitizer racecheck found exactly that; a perturbed run
// produced wrong rows).  Four slots, indexed by the tile counter & 3.
constexpr int HF_STATS_SLOTS = 4;
constexpr int HF_STATS_BYTES = HF_STATS_SLOTS * HF_MAXKB * 128 * 8;  // [tile & 3][key block][row] (max, sum)

struct FwdHrParams {
  HrGeom g;
  int nstage;
  int nkb;       // key blocks per head
  int kwa, kwl;  // stored keys (multiple of 16, <= 112) of the blocks jb < nkb - 1 / of the last block; block jb starts at jb * kwa
  int nbuf;      // score buffers of kwa TMEM columns each; the accumulators follow them
  bf16* o_tok;
  float* lse;
  unsigned long long* trace;
};

template <int N>
__device__ __forceinline__ uint32_t (&sub(uint32_t (&v)[HF_MAXCOLS], int i))[N] { return *reinterpret_cast<uint32_t(*)[N]>(&v[i]); }

// position of a unit in the CTA's sequence, advanced without divisions
struct HfUnit {
  int u, k, t, jb, buf, par;  // index, local head, query tile, key block, score buffer, phase parity of the buffer's barriers
  __device__ __forceinline__ void next(int nt, int nkb, int nbuf) {
    ++u;
    if (++jb == nkb) { jb = 0; if (++t == nt) { t = 0; ++k; } }
    if (++buf == nbuf) { buf = 0; par ^= 1; }
  }
};

__global__ void __launch_bounds__(HF_THREADS, 1)
attn_fwd_hr_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_qt,
                   const __grid_constant__ CUtensorMap tm_kt, const __grid_constant__ CUtensorMap tm_vt, FwdHrParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  const HrGeom& G = p.g;
  const int stage_bytes = 3 * G.tensor_bytes;
  float2* stats = reinterpret_cast<float2*>(smem + p.nstage * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stats) + HF_STATS_BYTES);
  uint64_t* full = bars;                 // [2] head operands landed
  uint64_t* empty = full + 2;            // [2] every MMA of the head has completed
  uint64_t* s_full = empty + 2;          // [HF_MAXBUF] S_buf ready (MMA -> the unit's softmax warpgroup)
  uint64_t* p_full = s_full + HF_MAXBUF; // [HF_MAXBUF] P written back over S_buf (4 warp arrivals)
  uint64_t* o_full = p_full + HF_MAXBUF; // every O_jb of the query tile accumulated
  uint64_t* o_empty = o_full + 1;        // merge warpgroup has read the O_jb (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = G.nt, L = G.L, nkb = p.nkb, nbuf = p.nbuf;
  const int n_local = (G.heads - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                      static_cast<int>(gridDim.x);
  const int U = nt * nkb;            // units per head
  const int n_units = n_local * U;
  Tracer tr(p.trace, warp);
  auto block_keys = [&](int jb) { return jb == nkb - 1 ? p.kwl : p.kwa; };   // stored keys of block jb
  const bool two_stages = p.nstage == 2;                                    // (no divisions on the issue paths)
  auto stage_of = [&](int k) { return two_stages ? (k & 1) : 0; };
  auto stage_par = [&](int k) { return two_stages ? ((k >> 1) & 1) : (k & 1); };

  // Rows of the tail tile beyond tail16 are never written by TMA but are read by the M = 128 MMAs: they must hold
  // finite values (their results are discarded), so the operand area starts out as zeros.
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = p.nstage * stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += HF_THREADS) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_qt); tma_prefetch_desc(&tm_kt); tma_prefetch_desc(&tm_vt);
    for (int s = 0; s < 2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < HF_MAXBUF; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 4); }
    mbar_init(o_full, 1);
    mbar_init(o_empty, 4);
    fence_mbar_init();
  }
  if (warp == 13) {
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
  // TMEM columns: S_i [i kwa, +kwa) for i < nbuf, then O_jb [nbuf kwa + 64 jb, +64).  P: bf16 pairs over the first half of S_i
  const uint32_t t_o = tmem_base + nbuf * p.kwa;

  if (warp < 8) {
    // ------------------------------------------------------------ softmax: warpgroup w owns the units u = w (mod 2)
    setmaxnreg_inc<168>();
    const int wg = warp >> 2, quad = warp & 3;
    const int row = quad * 32 + lane;
    HfUnit it{0, 0, 0, 0, 0, 0};
    if (wg == 1 && n_units > 1) it.next(nt, nkb, nbuf);
    for (; it.u < n_units; it.next(nt, nkb, nbuf), it.next(nt, nkb, nbuf)) {
      const int tc = it.k * nt + it.t, jb = it.jb;
      const int kw = block_keys(jb), nv = min(kw, L - jb * p.kwa);  // my columns / of which real keys
      const bool active = quad * 32 < hr_rows(G, it.t);  // warp-uniform: some row of this warp is a real query
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + it.buf * p.kwa;
      mbar_wait(&s_full[it.buf], it.par);
      tc_fence_after();
      tr(10);
      if (active) {
        uint32_t v[HF_MAXCOLS];
        if (kw >= 32) tmem_ld_32x32(t_row, sub<32>(v, 0)); else tmem_ld_32x16(t_row, sub<16>(v, 0));
        if (kw >= 64) tmem_ld_32x32(t_row + 32, sub<32>(v, 32)); else if (kw >= 48) tmem_ld_32x16(t_row + 32, sub<16>(v, 32));
        if (kw >= 96) tmem_ld_32x32(t_row + 64, sub<32>(v, 64)); else if (kw >= 80) tmem_ld_32x16(t_row + 64, sub<16>(v, 64));
        if (kw >= 112) tmem_ld_32x16(t_row + 96, sub<16>(v, 96));
        tmem_ld_wait();
        // Padding keys (only the last 16 columns of the head's last block can hold any) become -inf: they drop out of
        // the maximum and their probabilities are exact zeros, so the hot loops below need no masked variant (the
        // kernel is instruction-cache bound: every unrolled variant costs fetch stalls in all roles).
        if (nv < kw) {
#pragma unroll
          for (int c = 0; c < HF_MAXCOLS; c += 16) {
            if (c + 16 == kw) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c + i >= nv) v[c + i] = 0xff800000u;
            }
          }
        }
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < HF_MAXCOLS; c += 16) {
          if (c < kw) {
#pragma unroll
            for (int i = 0; i < 16; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[c + i]));
          }
        }
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        const float mxs = mx * LOG2E;
        tr(11);
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        // 32 scores -> 16 packed bf16 pairs, written back over the first half of the columns (all of them are in registers)
#pragma unroll
        for (int c = 0; c < HF_MAXCOLS; c += 32) {
          if (c < kw) {
            uint32_t pk[16];
#pragma unroll
            for (int h = 0; h < 32 && c + h < HF_MAXCOLS; h += 16) {
              if (c + h < kw) {
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                  const float e0 = fast_exp2(fmaf(__uint_as_float(v[c + h + i]), LOG2E, -mxs));
                  const float e1 = fast_exp2(fmaf(__uint_as_float(v[c + h + i + 1]), LOG2E, -mxs));
                  s4[(i >> 1) & 3] += e0 + e1;
                  pk[(h + i) >> 1] = pack_bf16(e0, e1);
                }
              }
            }
            if (c + 32 <= kw) tmem_st_32x16(t_row + (c >> 1), pk);
            else tmem_st_32x8(t_row + (c >> 1), *reinterpret_cast<const uint32_t(*)[8]>(&pk[0]));
          }
        }
        stats[((tc & (HF_STATS_SLOTS - 1)) * HF_MAXKB + jb) * 128 + row] = make_float2(mx, (s4[0] + s4[1]) + (s4[2] + s4[3]));
        tr(13);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[it.buf]);
      tr(14);
    }
  } else if (warp == 12) {
    setmaxnreg_dec<56>();
    // ------------------------------------------------------------ TMA producer: one head per stage
    if (elect_one()) {
      auto prefetch_head = [&](int k_) {
        if (k_ >= n_local) return;
        const int g_ = blockIdx.x + k_ * gridDim.x;
        for (int i = 0; i < nt - 1; ++i) {
          tma_prefetch_l2_3d(&tm_q, 0, i * 128, g_); tma_prefetch_l2_3d(&tm_k, 0, i * 128, g_);
          tma_prefetch_l2_3d(&tm_v, 0, i * 128, g_);
        }
        tma_prefetch_l2_3d(&tm_qt, 0, (nt - 1) * 128, g_); tma_prefetch_l2_3d(&tm_kt, 0, (nt - 1) * 128, g_);
        tma_prefetch_l2_3d(&tm_vt, 0, (nt - 1) * 128, g_);
      };
      // one head beyond the shared-memory stages: measured with 4 heads ahead, the ~60 MB of prefetched operands in
      // flight across 148 CTAs thrashed L2 (DRAM reads 1.8x the algorithmic bytes)
      for (int k = 0; k < n_local; ++k) {
        const int g = blockIdx.x + k * gridDim.x;
        const int s = stage_of(k);
        prefetch_head(k + p.nstage);
        mbar_wait(&empty[s], stage_par(k) ^ 1);
        tr(30);
        uint8_t* st = smem + s * stage_bytes;
        mbar_expect_tx(&full[s], static_cast<uint32_t>(stage_bytes));
        for (int i = 0; i < nt; ++i) {
          const bool tl = i == nt - 1;
          tma_load_3d(st + i * TILE_BYTES, tl ? &tm_qt : &tm_q, &full[s], 0, i * 128, g);
          tma_load_3d(st + G.tensor_bytes + i * TILE_BYTES, tl ? &tm_kt : &tm_k, &full[s], 0, i * 128, g);
          tma_load_3d(st + 2 * G.tensor_bytes + i * TILE_BYTES, tl ? &tm_vt : &tm_v, &full[s], 0, i * 128, g);
        }
      }
    }
  } else if (warp == 13) {
    setmaxnreg_dec<56>();
    // ------------------------------------------------------------ MMA issuer
    auto issue_s = [&](const HfUnit x) {
      const int s = stage_of(x.k);
      if (x.t == 0 && x.jb == 0) { mbar_wait(&full[s], stage_par(x.k)); tr(1); }
      tc_fence_after();
      tr(2);
      if (elect_one()) {
        const uint32_t st = smem_u32(smem + s * stage_bytes);
        const uint64_t dq = umma_desc_kmajor_sw128(st + x.t * TILE_BYTES);
        const uint64_t dk = umma_desc_kmajor_sw128(st + G.tensor_bytes + x.jb * p.kwa * 128);
        const uint32_t idesc = umma_idesc_bf16(128, block_keys(x.jb));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem_base + x.buf * p.kwa, dq + 2 * kk, dk + 2 * kk, idesc, kk != 0);
        umma_commit(&s_full[x.buf]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](const HfUnit x) {
      const int s = stage_of(x.k);
      const int tc = x.k * nt + x.t;
      mbar_wait(&p_full[x.buf], x.par);
      tr(3);
      if (x.jb == 0) { mbar_wait(o_empty, (tc & 1) ^ 1); tr(5); }  // the merge warpgroup has drained the previous tile's O_jb
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sv = smem_u32(smem + s * stage_bytes + 2 * G.tensor_bytes + x.jb * p.kwa * 128);
        constexpr uint32_t idesc = umma_idesc_bf16(128, 64) | IDESC_B_MN;
        const int ksteps = block_keys(x.jb) >> 4;
        const uint32_t pa = tmem_base + x.buf * p.kwa;
        for (int kk = 0; kk < ksteps; ++kk)  // A = P: 16 keys = 8 TMEM columns per step; B = V rows (64 d contiguous per key)
          umma_bf16_ts(t_o + x.jb * 64, pa + kk * 8, umma_desc_mnmajor_sw128(sv + kk * 2048, TILE_BYTES), idesc, kk != 0);
        if (x.jb == nkb - 1) umma_commit(o_full);
        if (x.t == nt - 1 && x.jb == nkb - 1) umma_commit(&empty[s]);
      }
      __syncwarp();
      tr(4);
    };
    // S_u goes into the buffer whose previous reader, P V of unit u - nbuf, has been issued (the tensor pipe runs in
    // order).  With a single operand stage the scores of the next head must not be issued before this head's last
    // P V: its operands only load once that product has completed.
    HfUnit xs{0, 0, 0, 0, 0, 0}, xp{0, 0, 0, 0, 0, 0};
    for (;;) {  // xp.u = number of P V products issued so far
      while (xs.u < n_units && xs.u < xp.u + nbuf && (p.nstage >= 2 || xs.k * U <= xp.u)) { issue_s(xs); xs.next(nt, nkb, nbuf); }
      if (xp.u >= n_units) break;
      issue_pv(xp);
      xp.next(nt, nkb, nbuf);
    }
  } else if (warp < 12) {
    setmaxnreg_dec<112>();
    // ------------------------------------------------------------ merge warpgroup: O = sum_jb w_jb O_jb / l
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_row = t_o + (static_cast<uint32_t>(quad * 32) << 16);
    int k = 0, t = 0;
    const int n_tiles = n_local * nt;
    for (int tc = 0; tc < n_tiles; ++tc) {
      const int g = blockIdx.x + k * gridDim.x;
      const bool active = quad * 32 < hr_rows(G, t);
      const int lq = t * 128 + row;
      mbar_wait(o_full, tc & 1);
      tc_fence_after();
      tr(20);
      if (active) {
        const float2* st = stats + (tc & (HF_STATS_SLOTS - 1)) * HF_MAXKB * 128 + row;  // [jb][row]
        float w[HF_MAXKB], m = -INFINITY, l = 0.f;
        float2 ms[HF_MAXKB];
#pragma unroll
        for (int j = 0; j < HF_MAXKB; ++j) {
          ms[j] = make_float2(-INFINITY, 0.f);
          if (j < nkb) { ms[j] = st[j * 128]; m = fmaxf(m, ms[j].x); }
        }
#pragma unroll
        for (int j = 0; j < HF_MAXKB; ++j) {
          w[j] = j < nkb ? fast_exp2((ms[j].x - m) * LOG2E) : 0.f;
          l = fmaf(w[j], ms[j].y, l);
        }
        const float inv = 1.f / l;
        const bool valid = lq < L;
        const int n = g / G.H, h = g - n * G.H;
        bf16* orow = p.o_tok + (static_cast<size_t>(valid ? lq : 0) * G.NB + n) * G.D + h * 64;
        float acc[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) acc[c] = 0.f;
#pragma unroll 1
        for (int j = 0; j < nkb; ++j) {  // not unrolled: code size (see above)
          const float wj = (j == 0 ? w[0] : j == 1 ? w[1] : j == 2 ? w[2] : w[3]) * inv;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            tmem_ld_32x32(t_row + j * 64 + half * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[half * 32 + c] = fmaf(wj, __uint_as_float(v[c]), acc[half * 32 + c]);
          }
        }
        // every O_jb of this tile is in registers: the next tile's P V products may overwrite them (before the slow,
        // row-strided global stores)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);
        if (valid) {
#pragma unroll
          for (int c = 0; c < 64; c += 8)
            *reinterpret_cast<uint4*>(orow + c) =
                make_uint4(pack_bf16(acc[c], acc[c + 1]), pack_bf16(acc[c + 2], acc[c + 3]),
                           pack_bf16(acc[c + 4], acc[c + 5]), pack_bf16(acc[c + 6], acc[c + 7]));
        }
        if (valid) p.lse[static_cast<size_t>(g) * L + lq] = m + __logf(l);
        tr(22);
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);
      }
      if (++t == nt) { t = 0; ++k; }
    }
  } else {
    setmaxnreg_dec<56>();  // warps 14, 15: complete the fourth warpgroup for setmaxnreg, no work
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ===================================================================================================== backward
// Operands Q', K, V', dO of the head stay resident (one stage; O visits the P / dS area at the start of the head, only
// to form delta = rowsum(dO o O)).  Per (query tile t, key block j) pair, non-transposed orientation:
//   MMA1  S = Q_t K_j^T -> TMEM [0,128),  dP = dO_t V_j^T -> TMEM [128,256)          (N = stored keys of block j)
//   WG0/1 P = exp(S - lse), dS = P o (dP - delta): the two warpgroups split the key columns (64 each), bf16 tiles in smem
//   MMA2  dV_j += P^T dO_t, dK_j += dS^T Q_t (K steps = stored rows of tile t), dQ_t += dS K_j (K steps = stored keys)
//   WG2   drains dV_j / dK_j after the last tile of the group and dQ_t after the last key block
// Query tiles are processed in groups of two (TMEM: 256 + 128 + 2 x 64 columns); a second group (L > 256) adds its
// dK / dV contribution onto the first group's output (same thread, same rows).
constexpr int HB_THREADS = 448;  // WG0, WG1 (P / dS), WG2 (gradients out), TMA warp, MMA warp
constexpr int HB_VEC_BYTES = 2 * 2 * 128 * HR_MAXT * 4;  // [head parity][delta | lse * log2e][384] fp32

struct BwdHrParams {
  HrGeom g;
  int ld;
  int nbox;      // staging boxes for the gradient stores (v2): 2 when shared memory allows, else 1
  int l2_prefetch;  // pull the next head's operands into L2 while this one computes (diagnostics: PEVIT_ATTN_BWD_PF=1; off by default)
  const float* lse;
  bf16* dqkv;
  bf16* ddelta;  // nullable
  unsigned long long* trace;
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

// W columns [c, c + W) of one query row: P and dS as bf16 into the swizzled [128][128] tiles
template <int W>
__device__ __forceinline__ void pds_chunk(uint32_t t_row, uint8_t* sP, uint8_t* sdS, int row, int c, int ncols_valid,
                                          bool row_valid, float lse_s, float delta) {
  uint32_t sv[W], dv[W];
  if constexpr (W == 32) { tmem_ld_32x32(t_row + c, sv); tmem_ld_32x32(t_row + 128 + c, dv); }
  else { tmem_ld_32x16(t_row + c, sv); tmem_ld_32x16(t_row + 128 + c, dv); }
  tmem_ld_wait();
#pragma unroll
  for (int q = 0; q < W / 8; ++q) {
    float pj[8], ds[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int col = c + 8 * q + i;
      const float e = (row_valid && col < ncols_valid) ? fast_exp2(fmaf(__uint_as_float(sv[8 * q + i]), LOG2E, -lse_s)) : 0.f;
      pj[i] = e;
      ds[i] = e * (__uint_as_float(dv[8 * q + i]) - delta);
    }
    const int kc = (c >> 3) + q;
    const int off = (kc >> 3) * TILE_BYTES + row * 128 + (((kc & 7) ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(sP + off) = make_uint4(pack_bf16(pj[0], pj[1]), pack_bf16(pj[2], pj[3]),
                                                     pack_bf16(pj[4], pj[5]), pack_bf16(pj[6], pj[7]));
    *reinterpret_cast<uint4*>(sdS + off) = make_uint4(pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]),
                                                      pack_bf16(ds[4], ds[5]), pack_bf16(ds[6], ds[7]));
  }
}

__global__ void __launch_bounds__(HB_THREADS, 1)
attn_bwd_hr_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                   const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_qt,
                   const __grid_constant__ CUtensorMap tm_kt, const __grid_constant__ CUtensorMap tm_vt,
                   const __grid_constant__ CUtensorMap tm_dot, const __grid_constant__ CUtensorMap tm_ot, BwdHrParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  const HrGeom& G = p.g;
  const int TB = G.tensor_bytes;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TB;
  uint8_t* sV = smem + 2 * TB;
  uint8_t* sdO = smem + 3 * TB;
  uint8_t* sP = smem + 4 * TB;       // [128 query rows][128 keys] bf16 as two [128][64] half tiles
  uint8_t* sdS = sP + 2 * TILE_BYTES;
  uint8_t* sO = sP;                  // O tiles of the head, only until delta is formed
  float* vecs = reinterpret_cast<float*>(sdS + 2 * TILE_BYTES);  // [parity][0: delta, 1: lse * log2e][384]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(vecs) + HB_VEC_BYTES);
  uint64_t* full = bars;               // head operands landed
  uint64_t* empty = full + 1;          // every MMA of the head has completed
  uint64_t* s_full = empty + 1;        // MMA1 (S, dP) of a pair done
  uint64_t* pds_full = s_full + 1;     // P, dS written (256 arrivals)
  uint64_t* mma2_done = pds_full + 1;  // MMA2 of a pair done reading P, dS
  uint64_t* kv_full = mma2_done + 1;   // dK_j, dV_j complete (all tiles of the group)
  uint64_t* kv_empty = kv_full + 1;    // WG2 drained dK_j, dV_j (128 arrivals)
  uint64_t* dq_full = kv_empty + 1;    // dQ of the group's tiles complete (all key blocks)
  uint64_t* dq_empty = dq_full + 1;    // WG2 drained dQ (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = G.nt, L = G.L;
  const int ngroups = (nt + 1) / 2;
  const int n_local = (G.heads - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                      static_cast<int>(gridDim.x);
  auto tiles_in_group = [&](int tg) { return min(2, nt - 2 * tg); };

  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = (4 * TB + 4 * TILE_BYTES) / 16;
    for (int i = threadIdx.x; i < n16; i += HB_THREADS) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_o); tma_prefetch_desc(&tm_qt); tma_prefetch_desc(&tm_kt); tma_prefetch_desc(&tm_vt);
    tma_prefetch_desc(&tm_dot); tma_prefetch_desc(&tm_ot);
    mbar_init(full, 1); mbar_init(empty, 1);
    mbar_init(s_full, 1); mbar_init(pds_full, 256); mbar_init(mma2_done, 1);
    mbar_init(kv_full, 1); mbar_init(kv_empty, 128); mbar_init(dq_full, 1); mbar_init(dq_empty, 128);
    fence_mbar_init();
  }
  if (warp == 13) {
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
  // TMEM columns: S [0,128) dP [128,256) dV_j [256,320) dK_j [320,384) dQ_tt [384 + 64 tt, +64)

  if (warp == 12) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      auto prefetch_head = [&](int k_) {  // pull the next head into L2 while this one computes
        if (k_ >= n_local) return;
        const int g_ = blockIdx.x + k_ * gridDim.x;
        const int n_ = g_ / G.H, h_ = g_ - n_ * G.H;
        for (int i = 0; i < nt; ++i) {
          const bool tl = i == nt - 1;
          tma_prefetch_l2_3d(tl ? &tm_qt : &tm_q, 0, i * 128, g_);
          tma_prefetch_l2_3d(tl ? &tm_kt : &tm_k, 0, i * 128, g_);
          tma_prefetch_l2_3d(tl ? &tm_vt : &tm_v, 0, i * 128, g_);
          tma_prefetch_l2_4d(tl ? &tm_dot : &tm_do, 0, h_, n_, i * 128);
          tma_prefetch_l2_4d(tl ? &tm_ot : &tm_o, 0, h_, n_, i * 128);
        }
      };
      for (int k = 0; k < n_local; ++k) {
        const int g = blockIdx.x + k * gridDim.x;
        const int n = g / G.H, h = g - n * G.H;
        prefetch_head(k + 1);  // one head ahead only: deeper prefetching thrashed L2 (see the forward kernel)
        mbar_wait(empty, (k & 1) ^ 1);
        mbar_expect_tx(full, static_cast<uint32_t>(5 * TB));
        for (int i = 0; i < nt; ++i) {
          const bool tl = i == nt - 1;
          tma_load_4d(sO + i * TILE_BYTES, tl ? &tm_ot : &tm_o, full, 0, h, n, i * 128);
          tma_load_4d(sdO + i * TILE_BYTES, tl ? &tm_dot : &tm_do, full, 0, h, n, i * 128);
          tma_load_3d(sQ + i * TILE_BYTES, tl ? &tm_qt : &tm_q, full, 0, i * 128, g);
          tma_load_3d(sK + i * TILE_BYTES, tl ? &tm_kt : &tm_k, full, 0, i * 128, g);
          tma_load_3d(sV + i * TILE_BYTES, tl ? &tm_vt : &tm_v, full, 0, i * 128, g);
        }
      }
    }
  } else if (warp == 13) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64) | IDESC_A_MN | IDESC_B_MN;  // A^T B forms
    constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64) | IDESC_B_MN;
    const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aDO = smem_u32(sdO);
    const uint32_t aP = smem_u32(sP), aS = smem_u32(sdS);
    auto issue_mma1 = [&](int t, int j) {
      if (elect_one()) {
        const uint32_t idesc_s = umma_idesc_bf16(128, hr_rows16(G, j));
        const uint64_t dq = umma_desc_kmajor_sw128(aQ + t * TILE_BYTES), dk = umma_desc_kmajor_sw128(aK + j * TILE_BYTES);
        const uint64_t dv = umma_desc_kmajor_sw128(aV + j * TILE_BYTES), ddo = umma_desc_kmajor_sw128(aDO + t * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem_base, dq + 2 * kk, dk + 2 * kk, idesc_s, kk != 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem_base + 128, ddo + 2 * kk, dv + 2 * kk, idesc_s, kk != 0);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    int bp = 0, kvc = 0, gc = 0;
    for (int k = 0; k < n_local; ++k) {
      mbar_wait(full, k & 1);
      tc_fence_after();
      issue_mma1(0, 0);
      for (int tg = 0; tg < ngroups; ++tg, ++gc) {
        const int ntg = tiles_in_group(tg);
        for (int j = 0; j < nt; ++j, ++kvc) {
          for (int tt = 0; tt < ntg; ++tt, ++bp) {
            const int t = 2 * tg + tt;
            mbar_wait(pds_full, bp & 1);
            if (tt == 0) mbar_wait(kv_empty, (kvc & 1) ^ 1);          // dK / dV accumulators drained (previous block)
            if (j == 0 && tt == 0) mbar_wait(dq_empty, (gc & 1) ^ 1);  // dQ accumulators drained (previous group)
            tc_fence_after();
            const bool last_tt = tt == ntg - 1;
            const bool last_of_head = last_tt && j == nt - 1 && tg == ngroups - 1;
            if (elect_one()) {
              const int qsteps = hr_rows16(G, t) >> 4, ksteps = hr_rows16(G, j) >> 4;
              for (int kk = 0; kk < qsteps; ++kk)  // dV_j (+)= P^T dO_t   (K = query rows, 16 per step)
                umma_bf16_ss(tmem_base + 256, umma_desc_mnmajor_sw128(aP + kk * 2048, TILE_BYTES),
                             umma_desc_mnmajor_sw128(aDO + t * TILE_BYTES + kk * 2048, TILE_BYTES), idesc_t, (tt | kk) != 0);
              for (int kk = 0; kk < qsteps; ++kk)  // dK_j (+)= dS^T Q_t
                umma_bf16_ss(tmem_base + 320, umma_desc_mnmajor_sw128(aS + kk * 2048, TILE_BYTES),
                             umma_desc_mnmajor_sw128(aQ + t * TILE_BYTES + kk * 2048, TILE_BYTES), idesc_t, (tt | kk) != 0);
              for (int kk = 0; kk < ksteps; ++kk)  // dQ_t (+)= dS K_j     (K = keys)
                umma_bf16_ss(tmem_base + 384 + tt * 64, umma_desc_kmajor_sw128(aS + (kk >> 2) * TILE_BYTES) + 2 * (kk & 3),
                             umma_desc_mnmajor_sw128(aK + j * TILE_BYTES + kk * 2048, TILE_BYTES), idesc_q, (j | kk) != 0);
              umma_commit(mma2_done);
              if (last_tt) umma_commit(kv_full);
              if (last_tt && j == nt - 1) umma_commit(dq_full);
              if (last_of_head) umma_commit(empty);
            }
            __syncwarp();
            if (!last_of_head) {  // S / dP of the next pair: both TMEM regions were consumed before pds_full
              int t2 = t + 1, j2 = j, tg2 = tg;
              if (tt == ntg - 1) { j2 = j + 1; t2 = 2 * tg; if (j2 == nt) { j2 = 0; tg2 = tg + 1; t2 = 2 * tg2; } }
              issue_mma1(t2, j2);
            }
          }
        }
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------ WG0 / WG1: delta, then P and dS of every pair
    const int wg = warp >> 2, quad = warp & 3;
    const int row = quad * 32 + lane;
    const int tid = threadIdx.x;  // 0..255
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    int bp = 0;
    for (int k = 0; k < n_local; ++k) {
      const int g = blockIdx.x + k * gridDim.x;
      float* vdelta = vecs + (k & 1) * 2 * 128 * HR_MAXT;
      float* vlse = vdelta + 128 * HR_MAXT;
      float lse_r[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int l = tid + i * 256;
        lse_r[i] = l < L ? p.lse[static_cast<size_t>(g) * L + l] * LOG2E : 0.f;
      }
      mbar_wait(full, k & 1);
      // delta[l] = sum_d dO[l][d] * O[l][d]: both tiles share the swizzle, so matching physical chunks pair up
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int l = tid + i * 256;
        if (l < L) {
          const uint8_t* pdo = sdO + (l >> 7) * TILE_BYTES + (l & 127) * 128;
          const uint8_t* po = sO + (l >> 7) * TILE_BYTES + (l & 127) * 128;
          float d = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int pc = (c ^ (l & 7)) * 16;  // a different physical chunk per row of a quarter-warp: no bank conflicts
            const uint4 a = *reinterpret_cast<const uint4*>(pdo + pc);
            const uint4 b = *reinterpret_cast<const uint4*>(po + pc);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 fa = unpack_bf16(aw[q]), fb = unpack_bf16(bw[q]);
              d = fmaf(fa.x, fb.x, fmaf(fa.y, fb.y, d));
            }
          }
          vdelta[l] = d;
          vlse[l] = lse_r[i];
        }
      }
      named_bar_sync(1, 256);  // delta / lse of the head visible to both warpgroups; the O tiles may now be overwritten
      for (int tg = 0; tg < ngroups; ++tg) {
        const int ntg = tiles_in_group(tg);
        for (int j = 0; j < nt; ++j) {
          const int kw = hr_rows16(G, j), kvalid = hr_rows(G, j);
          for (int tt = 0; tt < ntg; ++tt, ++bp) {
            const int t = 2 * tg + tt;
            const int lq = t * 128 + row;
            const bool active = quad * 32 < hr_rows16(G, t);  // warp holds rows the MMAs read (real or zero padding)
            const bool row_valid = row < hr_rows(G, t);
            const float lse_s = row_valid ? vlse[lq] : 0.f;
            const float delta = row_valid ? vdelta[lq] : 0.f;
            mbar_wait(s_full, bp & 1);
            tc_fence_after();
            if (bp > 0) mbar_wait(mma2_done, (bp - 1) & 1);  // MMA2 of the previous pair is done reading P / dS
            if (active) {
              for (int c = wg * 64; c < min(kw, wg * 64 + 64); c += 32) {
                if (c + 32 <= kw) pds_chunk<32>(t_row, sP, sdS, row, c, kvalid, row_valid, lse_s, delta);
                else pds_chunk<16>(t_row, sP, sdS, row, c, kvalid, row_valid, lse_s, delta);
              }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(pds_full);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ WG2: gradients out
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const size_t plane = static_cast<size_t>(G.heads) * L * 64;
    int kvc = 0, gc = 0;
    for (int k = 0; k < n_local; ++k) {
      const int g = blockIdx.x + k * gridDim.x;
      const int n = g / G.H, h = g - n * G.H;
      for (int tg = 0; tg < ngroups; ++tg, ++gc) {
        const int ntg = tiles_in_group(tg);
        const bool add_prev = tg > 0;  // a previous group of this head already wrote its dK / dV share
        for (int j = 0; j < nt; ++j, ++kvc) {
          const int lk = j * 128 + row;
          const bool valid = row < hr_rows(G, j);
          const bool active = quad * 32 < hr_rows(G, j);
          bf16* tok = p.dqkv + (static_cast<size_t>(valid ? lk : 0) * G.NB + n) * p.ld + h * 64;
          bf16* hm = p.ddelta != nullptr ? p.ddelta + plane + (static_cast<size_t>(g) * L + (valid ? lk : 0)) * 64 : nullptr;
          mbar_wait(kv_full, kvc & 1);
          tc_fence_after();
          if (active) {
#pragma unroll 1
            for (int part = 0; part < 2; ++part) {  // 0: dV', 1: dK
              bf16* dst_tok = tok + (part == 0 ? 2 * G.D : G.D);
#pragma unroll
              for (int c = 0; c < 64; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + 256 + part * 64 + c, v);
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
          }
          tc_fence_before();
          mbar_arrive(kv_empty);
        }
        mbar_wait(dq_full, gc & 1);
        tc_fence_after();
        for (int tt = 0; tt < ntg; ++tt) {
          const int t = 2 * tg + tt;
          const int lq = t * 128 + row;
          const bool valid = row < hr_rows(G, t);
          if (quad * 32 < hr_rows(G, t)) {
            bf16* dst_tok = p.dqkv + (static_cast<size_t>(valid ? lq : 0) * G.NB + n) * p.ld + h * 64;
            bf16* dst_hm = p.ddelta != nullptr ? p.ddelta + (static_cast<size_t>(g) * L + (valid ? lq : 0)) * 64 : nullptr;
#pragma unroll
            for (int c = 0; c < 64; c += 32) {
              uint32_t v[32];
              tmem_ld_32x32(t_row + 384 + tt * 64 + c, v);
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
        }
        tc_fence_before();
        mbar_arrive(dq_empty);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ===================================================================================================== backward, v2
// Same residency and the same accumulators as above, but in the TRANSPOSED orientation and in half-tile units, so that
// the exponentials of one unit overlap the tensor work of the previous one and the probabilities never visit shared
// memory:
//   unit  = (query tile t, 64-row half h, key block j).  TMEM buffer b = unit & 1 (two of them: 2 x 128 columns):
//   MMA1  S^T = K_j Q_{t,h}^T -> b*128 + [0,64),  dP^T = V_j dO_{t,h}^T -> b*128 + [64,128)   (M = 128 keys, N = <= 64 rows)
//   WG0/1 one thread per KEY lane, the warpgroups split the 64 query columns: P^T = exp(S^T - lse_q),
//         dS^T = P^T o (dP^T - delta_q); both written back into TMEM over the columns just read (bf16 pairs), dS^T also
//         into a [128 keys][128 rows] shared-memory tile
//   MMA2  dV_j += P^T dO_{t,h}, dK_j += dS^T Q_{t,h}: tcgen05.mma with the A operand in TMEM; after the tile's last half
//         dQ_t += dS_t K_j from the shared-memory tile (A read MN-major)
// While the warpgroups work on unit u+1 (TMEM-read bound: 64 KB per unit at 64 B/clk), the tensor pipe runs MMA2(u) and
// MMA1(u+2); v1 above serialises the two phases.
struct HrUnit { int t, j, tt, h, nq, flags, pair, pad; };
enum { U_FIRST_KV = 1, U_LAST_KV = 2, U_FIRST_GRP = 4, U_LAST_GRP = 8, U_LAST_HALF = 16, U_LAST = 32 };
constexpr int HB2_MAX_UNITS = 24;

__global__ void __launch_bounds__(HB_THREADS, 1)
attn_bwd_hr2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                    const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_qt,
                    const __grid_constant__ CUtensorMap tm_kt, const __grid_constant__ CUtensorMap tm_vt,
                    const __grid_constant__ CUtensorMap tm_dot, const __grid_constant__ CUtensorMap tm_ot,
                    const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_gt,
                    const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_ht, BwdHrParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  const HrGeom& G = p.g;
  const int TB = G.tensor_bytes;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TB;
  uint8_t* sV = smem + 2 * TB;
  uint8_t* sdO = smem + 3 * TB;
  uint8_t* sDS = smem + 4 * TB;      // [pair parity][q half][128 key rows][64 query cols] bf16: dS^T of the current pairs
  uint8_t* sO = sDS;                 // O tiles of the head, only until delta is formed
  uint8_t* stg = sDS + 4 * TILE_BYTES;  // nbox x [128 rows][64] bf16 boxes the gradients leave through (TMA store)
  float* vecs = reinterpret_cast<float*>(stg + p.nbox * TILE_BYTES);  // [parity][0: delta, 1: lse * log2e][384]
  HrUnit* units = reinterpret_cast<HrUnit*>(reinterpret_cast<uint8_t*>(vecs) + HB_VEC_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(units + HB2_MAX_UNITS);
  uint64_t* full = bars;               // head operands landed
  uint64_t* empty = full + 1;          // every MMA of the head has completed
  uint64_t* s_full = empty + 1;        // [2] MMA1 of a unit done
  uint64_t* pds_full = s_full + 2;     // [2] P^T, dS^T of a unit written (8 warp arrivals)
  uint64_t* ds_free = pds_full + 2;    // [2] dQ MMA of a pair done reading its dS^T tile
  uint64_t* kv_full = ds_free + 2;     // dK_j, dV_j complete (all tiles of the group)
  uint64_t* kv_empty = kv_full + 1;    // WG2 drained dK_j, dV_j (4 warp arrivals)
  uint64_t* dq_full = kv_empty + 1;    // dQ of the group's tiles complete (all key blocks)
  uint64_t* dq_empty = dq_full + 1;    // WG2 drained dQ (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_empty + 1);
  int* n_units_s = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = G.nt, L = G.L;
  const int ngroups = (nt + 1) / 2;
  const int n_local = (G.heads - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                      static_cast<int>(gridDim.x);
  auto tiles_in_group = [&](int tg) { return min(2, nt - 2 * tg); };
  Tracer tr(p.trace, warp);

  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = (4 * TB + 4 * TILE_BYTES) / 16;
    for (int i = threadIdx.x; i < n16; i += HB_THREADS) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {  // the unit sequence of one head (identical for every head)
    int n = 0, pair = 0;
    for (int tg = 0; tg < ngroups; ++tg) {
      const int ntg = tiles_in_group(tg);
      for (int j = 0; j < nt; ++j) {
        for (int tt = 0; tt < ntg; ++tt, ++pair) {
          const int t = 2 * tg + tt;
          const int r16 = hr_rows16(G, t);
          const int nh = r16 > 64 ? 2 : 1;
          for (int h = 0; h < nh; ++h, ++n) {
            HrUnit u{t, j, tt, h, min(64, r16 - 64 * h), 0, pair, 0};
            if (tt == 0 && h == 0) u.flags |= U_FIRST_KV;
            if (tt == ntg - 1 && h == nh - 1) u.flags |= U_LAST_KV;
            if (j == 0 && tt == 0 && h == 0) u.flags |= U_FIRST_GRP;
            if (j == nt - 1 && tt == ntg - 1 && h == nh - 1) u.flags |= U_LAST_GRP;
            if (h == nh - 1) u.flags |= U_LAST_HALF;
            if (tg == ngroups - 1 && j == nt - 1 && tt == ntg - 1 && h == nh - 1) u.flags |= U_LAST;
            units[n] = u;
          }
        }
      }
    }
    *n_units_s = n;
  }
  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_o); tma_prefetch_desc(&tm_qt); tma_prefetch_desc(&tm_kt); tma_prefetch_desc(&tm_vt);
    tma_prefetch_desc(&tm_dot); tma_prefetch_desc(&tm_ot);
    tma_prefetch_desc(&tm_g); tma_prefetch_desc(&tm_gt); tma_prefetch_desc(&tm_h); tma_prefetch_desc(&tm_ht);
    mbar_init(full, 1); mbar_init(empty, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&pds_full[b], 8); mbar_init(&ds_free[b], 1); }
    mbar_init(kv_full, 1); mbar_init(kv_empty, 4); mbar_init(dq_full, 1); mbar_init(dq_empty, 4);
    fence_mbar_init();
  }
  if (warp == 13) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_units = *n_units_s;
  const int pairs_per_head = units[n_units - 1].pair + 1;
  pdl_launch_dependents();
  pdl_wait();
  // TMEM columns: buffer b: S^T [128 b, +64), dP^T [128 b + 64, +64);  dV_j [256,320) dK_j [320,384) dQ_tt [384 + 64 tt, +64)

  if (warp == 12) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      auto prefetch_head = [&](int k_) {
        if (k_ >= n_local || !p.l2_prefetch) return;
        const int g_ = blockIdx.x + k_ * gridDim.x;
        const int n_ = g_ / G.H, h_ = g_ - n_ * G.H;
        for (int i = 0; i < nt; ++i) {
          const bool tl = i == nt - 1;
          tma_prefetch_l2_3d(tl ? &tm_qt : &tm_q, 0, i * 128, g_);
          tma_prefetch_l2_3d(tl ? &tm_kt : &tm_k, 0, i * 128, g_);
          tma_prefetch_l2_3d(tl ? &tm_vt : &tm_v, 0, i * 128, g_);
          tma_prefetch_l2_4d(tl ? &tm_dot : &tm_do, 0, h_, n_, i * 128);
          tma_prefetch_l2_4d(tl ? &tm_ot : &tm_o, 0, h_, n_, i * 128);
        }
      };
      for (int k = 0; k < n_local; ++k) {
        const int g = blockIdx.x + k * gridDim.x;
        const int n = g / G.H, h = g - n * G.H;
        prefetch_head(k + 1);
        mbar_wait(empty, (k & 1) ^ 1);
        tr(30);
        mbar_expect_tx(full, static_cast<uint32_t>(5 * TB));
        for (int i = 0; i < nt; ++i) {
          const bool tl = i == nt - 1;
          tma_load_4d(sO + i * TILE_BYTES, tl ? &tm_ot : &tm_o, full, 0, h, n, i * 128);
          tma_load_4d(sdO + i * TILE_BYTES, tl ? &tm_dot : &tm_do, full, 0, h, n, i * 128);
          tma_load_3d(sQ + i * TILE_BYTES, tl ? &tm_qt : &tm_q, full, 0, i * 128, g);
          tma_load_3d(sK + i * TILE_BYTES, tl ? &tm_kt : &tm_k, full, 0, i * 128, g);
          tma_load_3d(sV + i * TILE_BYTES, tl ? &tm_vt : &tm_v, full, 0, i * 128, g);
        }
      }
    }
  } else if (warp == 13) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_ts = umma_idesc_bf16(128, 64) | IDESC_B_MN;               // A in TMEM, B = [rows][64 d]
    constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 64) | IDESC_A_MN | IDESC_B_MN;  // A = dS^T tile read transposed
    const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aDO = smem_u32(sdO), aDS = smem_u32(sDS);
    long long ug = 0;  // units issued (MMA1) so far, over all heads
    auto issue_mma1 = [&](const HrUnit& u, long long idx) {
      const int b = static_cast<int>(idx & 1);
      if (elect_one()) {
        const uint32_t idesc = umma_idesc_bf16(128, u.nq);
        const uint32_t roff = u.t * TILE_BYTES + u.h * 64 * 128;
        const uint64_t dk = umma_desc_kmajor_sw128(aK + u.j * TILE_BYTES), dv = umma_desc_kmajor_sw128(aV + u.j * TILE_BYTES);
        const uint64_t dq = umma_desc_kmajor_sw128(aQ + roff), ddo = umma_desc_kmajor_sw128(aDO + roff);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem_base + b * 128, dk + 2 * kk, dq + 2 * kk, idesc, kk != 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem_base + b * 128 + 64, dv + 2 * kk, ddo + 2 * kk, idesc, kk != 0);
        umma_commit(&s_full[b]);
      }
      __syncwarp();
    };
    long long uc = 0;   // units completed (MMA2 issued)
    int kvc = 0, gc = 0;
    for (int k = 0; k < n_local; ++k) {
      mbar_wait(full, k & 1);
      tc_fence_after();
      tr(1);
      issue_mma1(units[0], ug++);
      tr(46);
      if (n_units > 1) issue_mma1(units[1], ug++);
      tr(47);
      for (int ui = 0; ui < n_units; ++ui, ++uc) {
        const HrUnit u = units[ui];
        const int b = static_cast<int>(uc & 1);
        const long long pg = static_cast<long long>(k) * pairs_per_head + u.pair;  // pair counter over all heads
        mbar_wait(&pds_full[b], static_cast<uint32_t>((uc >> 1) & 1));
        tr(3);
        if (u.flags & U_FIRST_KV) { mbar_wait(kv_empty, (kvc & 1) ^ 1); tr(5); }   // dK / dV accumulators drained (previous block)
        if (u.flags & U_FIRST_GRP) { mbar_wait(dq_empty, (gc & 1) ^ 1); tr(6); }   // dQ accumulators drained (previous group)
        tc_fence_after();
        tr(42);
        if (elect_one()) {
          const uint32_t rows = u.t * TILE_BYTES + u.h * 64 * 128;  // first row of this half inside dO / Q
          const int qsteps = u.nq >> 4;
          for (int kk = 0; kk < qsteps; ++kk)   // dV_j (+)= P^T dO_{t,h}   (K = query rows of the half, 16 per step)
            umma_bf16_ts(tmem_base + 256, tmem_base + b * 128 + (kk >> 1) * 32 + (kk & 1) * 8,
                         umma_desc_mnmajor_sw128(aDO + rows + kk * 2048, TILE_BYTES), idesc_ts,
                         (!(u.flags & U_FIRST_KV) || kk != 0) ? 1u : 0u);
          tr(43);
          {                                     // dK_j (+)= dS^T Q_{t,h}: A = the shared-memory dS^T half tile (K-major: 64
            // query columns per key row).  From TMEM the A operand of an N = 64 product costs 64 cycles (4 KB at 64 B/clk),
            // twice the math; shared memory delivers it at 128 B/clk, and the warpgroups no longer write dS^T back to TMEM.
            const uint64_t dds = umma_desc_kmajor_sw128(aDS + static_cast<uint32_t>(pg & 1) * 2 * TILE_BYTES + u.h * TILE_BYTES);
            for (int kk = 0; kk < qsteps; ++kk)
              umma_bf16_ss(tmem_base + 320, dds + 2 * kk, umma_desc_mnmajor_sw128(aQ + rows + kk * 2048, TILE_BYTES), idesc_ts,
                           (!(u.flags & U_FIRST_KV) || kk != 0) ? 1u : 0u);
          }
          tr(44);
          if (u.flags & U_LAST_HALF) {          // dQ_t (+)= dS_t K_j       (K = stored keys of block j)
            const uint32_t ds = aDS + static_cast<uint32_t>(pg & 1) * 2 * TILE_BYTES;
            const int ksteps = hr_rows16(G, u.j) >> 4;
            for (int kk = 0; kk < ksteps; ++kk)
              umma_bf16_ss(tmem_base + 384 + u.tt * 64, umma_desc_mnmajor_sw128(ds + kk * 2048, TILE_BYTES),
                           umma_desc_mnmajor_sw128(aK + u.j * TILE_BYTES + kk * 2048, TILE_BYTES), idesc_dq,
                           (u.j | kk) != 0 ? 1u : 0u);
            tr(45);
            umma_commit(&ds_free[pg & 1]);
          }
          if (u.flags & U_LAST_KV) umma_commit(kv_full);
          if (u.flags & U_LAST_GRP) umma_commit(dq_full);
          if (u.flags & U_LAST) umma_commit(empty);
        }
        __syncwarp();
        if (u.flags & U_LAST_KV) ++kvc;
        if (u.flags & U_LAST_GRP) ++gc;
        tr(4);
        if (ui + 2 < n_units) issue_mma1(units[ui + 2], ug++);  // into the buffer this unit has just released (in order)
        tr(2);
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------ WG0 / WG1: delta, then P^T and dS^T of every unit
    const int wg = warp >> 2, quad = warp & 3;
    const int row = quad * 32 + lane;  // key lane
    const int tid = threadIdx.x;       // 0..255
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    long long uc = 0;
    for (int k = 0; k < n_local; ++k) {
      const int g = blockIdx.x + k * gridDim.x;
      float* vdelta = vecs + (k & 1) * 2 * 128 * HR_MAXT;
      float* vlse = vdelta + 128 * HR_MAXT;
      float lse_r[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int l = tid + i * 256;
        lse_r[i] = l < L ? p.lse[static_cast<size_t>(g) * L + l] * LOG2E : 0.f;
      }
      mbar_wait(full, k & 1);
      tr(9);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int l = tid + i * 256;
        if (l < 128 * HR_MAXT) {
          float d = 0.f;
          if (l < L) {
            const uint8_t* pdo = sdO + (l >> 7) * TILE_BYTES + (l & 127) * 128;
            const uint8_t* po = sO + (l >> 7) * TILE_BYTES + (l & 127) * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int pc = (c ^ (l & 7)) * 16;  // a different physical chunk per row of a quarter-warp: reading chunk c
                                                  // of 8 consecutive rows is an 8-way bank conflict (4.7 k cycles per head)
              const uint4 a = *reinterpret_cast<const uint4*>(pdo + pc);
              const uint4 bq = *reinterpret_cast<const uint4*>(po + pc);
              const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 fa = unpack_bf16(aw[q]), fb = unpack_bf16(bw[q]);
                d = fmaf(fa.x, fb.x, fmaf(fa.y, fb.y, d));
              }
            }
          }
          vdelta[l] = d;                       // rows at or beyond L: zeros (their probabilities are masked anyway)
          vlse[l] = l < L ? lse_r[i] : 0.f;
        }
      }
      tr(16);
      named_bar_sync(1, 256);  // delta / lse of the head visible to both warpgroups; the O tiles may now be overwritten
      tr(17);
      for (int ui = 0; ui < n_units; ++ui, ++uc) {
        const HrUnit u = units[ui];
        const int b = static_cast<int>(uc & 1);
        const long long pg = static_cast<long long>(k) * pairs_per_head + u.pair;
        const int kw = hr_rows16(G, u.j), kvalid = hr_rows(G, u.j);
        const bool active = quad * 32 < kw;       // warp holds key lanes the dQ product reads (real or zero padding)
        const bool key_valid = row < kvalid;
        const int q0 = u.h * 64 + wg * 32;         // first query row (within the tile) of this thread's columns
        const int ncols = max(0, min(32, u.nq - wg * 32));
        const int qvalid = hr_rows(G, u.t) - q0;   // columns < qvalid are real query rows
        mbar_wait(&s_full[b], static_cast<uint32_t>((uc >> 1) & 1));
        tc_fence_after();
        tr(10);
        // the dS^T tile of this pair was last read by the dQ product two pairs ago
        if (u.h == 0) { mbar_wait(&ds_free[pg & 1], static_cast<uint32_t>(((pg >> 1) & 1) ^ 1)); tr(15); }
        // (measured: giving each warpgroup its own unit -- all 64 columns, no shared unit -- is SLOWER, 581 vs 563 us at
        // L = 197: the MMA warp issues ~20 instructions per unit at ~65 cycles each (A-operand fetch bound), so a
        // warpgroup that owns a buffer alone waits ~1800 cycles for its next scores; sharing every unit halves that.)
        if (active && ncols > 0) {
          const uint32_t ts = t_lane + b * 128 + wg * 32, tdp = ts + 64;
          uint32_t sv[32], dv[32];
          if (ncols == 32) { tmem_ld_32x32(ts, sv); tmem_ld_32x32(tdp, dv); }
          else { tmem_ld_32x16(ts, *reinterpret_cast<uint32_t(*)[16]>(&sv[0])); tmem_ld_32x16(tdp, *reinterpret_cast<uint32_t(*)[16]>(&dv[0])); }
          tmem_ld_wait();
          tr(11);
          const float* pl = vlse + u.t * 128 + q0;
          const float* pd = vdelta + u.t * 128 + q0;
          uint32_t pp[16], pds[16];
          uint8_t* srow = sDS + static_cast<uint32_t>(pg & 1) * 2 * TILE_BYTES + u.h * TILE_BYTES + row * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q * 8 < ncols) {
              const float4 l0 = *reinterpret_cast<const float4*>(pl + 8 * q), l1 = *reinterpret_cast<const float4*>(pl + 8 * q + 4);
              const float4 d0 = *reinterpret_cast<const float4*>(pd + 8 * q), d1 = *reinterpret_cast<const float4*>(pd + 8 * q + 4);
              const float ls[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
              const float dl[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
              float pj[8], ds[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int c = 8 * q + i;
                const float e = (key_valid && c < qvalid) ? fast_exp2(fmaf(__uint_as_float(sv[c]), LOG2E, -ls[i])) : 0.f;
                pj[i] = e;
                ds[i] = e * (__uint_as_float(dv[c]) - dl[i]);
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                pp[4 * q + i] = pack_bf16(pj[2 * i], pj[2 * i + 1]);
                pds[4 * q + i] = pack_bf16(ds[2 * i], ds[2 * i + 1]);
              }
              const int ch = wg * 4 + q;  // 16-byte chunk of the 64 query columns of this half
              *reinterpret_cast<uint4*>(srow + ((ch ^ (row & 7)) << 4)) = make_uint4(pds[4 * q], pds[4 * q + 1], pds[4 * q + 2], pds[4 * q + 3]);
            }
          }
          if (ncols == 32) tmem_st_32x16(ts, pp);   // P^T stays in TMEM (A operand of dV); dS^T is read from shared memory
          else tmem_st_32x8(ts, *reinterpret_cast<const uint32_t(*)[8]>(&pp[0]));
          tr(13);
          tmem_st_wait();
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pds_full[b]);
        tr(14);
      }
    }
  } else {
    // ------------------------------------------------------------ WG2: gradients out (same accumulators as v1)
    // Every gradient tile leaves as bulk TMA stores out of ONE swizzled [rows][64] staging box: token-major into dqkv
    // (4-D map: rows l * NB + n, head column block) and, for dV' / dQ', head-major into d(delta).  Per-thread 16-byte
    // stores of row-strided data cost 32 LSU wavefronts per instruction and kept the MIO queue of every sub-partition
    // busy for ~16 k cycles per head (and the next head's first MMAs waiting behind them).
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int r7 = row & 7;
    const bool has_hm = p.ddelta != nullptr;
    const bool two_boxes = p.nbox == 2;
    uint8_t* box = stg;                 // the box being filled / stored (alternates when there are two)
    uint8_t* rp = box + row * 128;
    auto leader = [&]() { return warp == 8 && elect_one(); };
    // TMEM columns [col, col + 64) of this thread's lane -> its row of the staging box, scaled
    auto stage_rows = [&](uint32_t col, float sc, bool active) {
      if (two_boxes) {                          // the stores out of THIS box (two rounds ago) have read it
        box = box == stg ? stg + TILE_BYTES : stg;
        rp = box + row * 128;
        if (leader()) tma_store_wait_read<1>();
      } else if (leader()) {
        tma_store_wait_read<0>();               // the previous stores have read the box
      }
      named_bar_sync(3, 128);
      if (active) {
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + col + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(rp + ((((c >> 3) + q) ^ r7) << 4)) =
                make_uint4(pack_bf16(__uint_as_float(v[8 * q]) * sc, __uint_as_float(v[8 * q + 1]) * sc),
                           pack_bf16(__uint_as_float(v[8 * q + 2]) * sc, __uint_as_float(v[8 * q + 3]) * sc),
                           pack_bf16(__uint_as_float(v[8 * q + 4]) * sc, __uint_as_float(v[8 * q + 5]) * sc),
                           pack_bf16(__uint_as_float(v[8 * q + 6]) * sc, __uint_as_float(v[8 * q + 7]) * sc));
        }
      }
    };
    int kvc = 0, gc = 0;
    for (int k = 0; k < n_local; ++k) {
      const int g = blockIdx.x + k * gridDim.x;
      const int n = g / G.H, h = g - n * G.H;
      for (int tg = 0; tg < ngroups; ++tg, ++gc) {
        const int ntg = tiles_in_group(tg);
        const bool add_prev = tg > 0;  // a previous group of this head already stored its dK / dV share: reduce-add
        for (int j = 0; j < nt; ++j, ++kvc) {
          const bool tail = j == nt - 1;
          const bool active = quad * 32 < hr_rows16(G, j);
          mbar_wait(kv_full, kvc & 1);
          tc_fence_after();
          tr(20);
#pragma unroll 1
          for (int part = 0; part < 2; ++part) {  // 0: dV', 1: dK
            stage_rows(256 + part * 64, 1.f, active);
            if (part == 1) {  // both accumulators are in registers / the box: the next key block may overwrite them
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(kv_empty);
            }
            fence_proxy_async_smem();
            named_bar_sync(3, 128);
            if (leader()) {
              const int hb = (part == 0 ? 2 * G.H : G.H) + h;  // dqkv column blocks: [dq | dk | dv] heads
              if (add_prev) {
                tma_store_wait_all<0>();  // the first group's stores of these rows have landed
                tma_reduce_add_4d(tail ? &tm_gt : &tm_g, box, 0, hb, n, j * 128);
                if (part == 0 && has_hm) tma_reduce_add_3d(tail ? &tm_ht : &tm_h, box, 0, j * 128, G.heads + g);
              } else {
                tma_store_4d(tail ? &tm_gt : &tm_g, box, 0, hb, n, j * 128);
                if (part == 0 && has_hm) tma_store_3d(tail ? &tm_ht : &tm_h, box, 0, j * 128, G.heads + g);
              }
              tma_store_commit();
            }
          }
          tr(21);
        }
        mbar_wait(dq_full, gc & 1);
        tc_fence_after();
        tr(23);
        for (int tt = 0; tt < ntg; ++tt) {
          const int t = 2 * tg + tt;
          const bool tail = t == nt - 1;
          const bool active = quad * 32 < hr_rows16(G, t);
          // token-major dQ = dQ' / 8 (q' = q / 8), head-major d(delta_q) = dQ': two passes over the same accumulator
          for (int pass = 0; pass < (has_hm ? 2 : 1); ++pass) {
            stage_rows(384 + tt * 64, pass == 0 ? 0.125f : 1.f, active);
            if (tt == ntg - 1 && pass == (has_hm ? 1 : 0)) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(dq_empty);
            }
            fence_proxy_async_smem();
            named_bar_sync(3, 128);
            if (leader()) {
              if (pass == 0) tma_store_4d(tail ? &tm_gt : &tm_g, box, 0, h, n, t * 128);
              else tma_store_3d(tail ? &tm_ht : &tm_h, box, 0, t * 128, g);
              tma_store_commit();
            }
          }
        }
        tr(24);
      }
    }
    if (leader()) tma_store_wait_all<0>();  // shared memory must outlive the last stores
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool attn_hr_supported(const AttnShape& a) { return a.r == 0 && a.L > 128 && a.L <= 128 * HR_MAXT && a.H * 64 == a.D; }

int attn_fwd_hr(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, bf16* o_tok, float* lse) {
  PEVIT_REQUIRE(attn_hr_supported(a), "attn_fwd_hr: unsupported shape L=%d D=%d H=%d r=%d", a.L, a.D, a.H, a.r);
  const HrGeom g = make_geom(a);
  CUtensorMap tq, tk, tv, tqt, tkt, tvt;
  if (make_tmap_bf16_hm3d(&tq, q, a.L, g.heads, 128) != 0) return -1;
  if (make_tmap_bf16_hm3d(&tk, k, a.L, g.heads, 128) != 0) return -1;
  if (make_tmap_bf16_hm3d(&tv, v, a.L, g.heads, 128) != 0) return -1;
  if (make_tmap_bf16_hm3d(&tqt, q, a.L, g.heads, g.tail16) != 0) return -1;
  if (make_tmap_bf16_hm3d(&tkt, k, a.L, g.heads, g.tail16) != 0) return -1;
  if (make_tmap_bf16_hm3d(&tvt, v, a.L, g.heads, g.tail16) != 0) return -1;
  const int stage_bytes = 3 * g.tensor_bytes;
  const int fixed = HF_STATS_BYTES + 256 + 1024;
  const int nstage = (227 * 1024 - fixed) / stage_bytes >= 2 ? 2 : 1;
  const int smem_bytes = nstage * stage_bytes + fixed;
  TraceHost trace;
  FwdHrParams p{};
  p.g = g; p.nstage = nstage; p.o_tok = o_tok; p.lse = lse;
  const int l16 = (g.nt - 1) * 128 + g.tail16;  // stored keys of a head
  p.nkb = (l16 + HF_MAXCOLS - 1) / HF_MAXCOLS;
  p.kwa = p.nkb == 1 ? l16 : ((l16 + p.nkb - 1) / p.nkb + 15) / 16 * 16;
  p.kwl = l16 - (p.nkb - 1) * p.kwa;
  p.nbuf = (512 - 64 * p.nkb) / p.kwa < HF_MAXBUF ? (512 - 64 * p.nkb) / p.kwa : HF_MAXBUF;
  PEVIT_REQUIRE(p.nkb <= HF_MAXKB && p.kwl >= 16 && p.kwl <= p.kwa && p.kwa <= HF_MAXCOLS && p.nbuf >= 2,
                "attn_fwd_hr: L=%d does not fit (key blocks %d x %d + %d, %d score buffers)", a.L, p.nkb - 1, p.kwa, p.kwl, p.nbuf);
  p.trace = trace.begin(HF_THREADS / 32);
  const int grid = g.heads < sm_count() ? g.heads : sm_count();
  static int configured[64] = {};
  int dev = 0;
  PEVIT_CHECK_CUDA(cudaGetDevice(&dev));
  if (configured[dev & 63] < smem_bytes) {
    PEVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_hr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[dev & 63] = 227 * 1024;
  }
  ProfScope prof(s, PC_ATTN_FWD);
  PEVIT_CHECK_CUDA(launch_kernel(attn_fwd_hr_kernel, dim3(grid), dim3(HF_THREADS), smem_bytes, s, 1, tq, tk, tv, tqt, tkt, tvt, p));
  PEVIT_CHECK_LAUNCH();
  trace.end(s, "fwd");
  return 0;
}

bool attn_bwd_hr_supported(const AttnShape& a) {
  if (!attn_hr_supported(a)) return false;
  const HrGeom g = make_geom(a);
  return 4 * g.tensor_bytes + 5 * TILE_BYTES + HB_VEC_BYTES + HB2_MAX_UNITS * static_cast<int>(sizeof(HrUnit)) + 256 + 1024 <= 227 * 1024 &&
         a.L <= 384;
}

int attn_bwd_hr(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, const bf16* o_tok,
                const bf16* do_tok, const float* lse, bf16* dqkv, int ld_dqkv, bf16* ddelta) {
  PEVIT_REQUIRE(attn_bwd_hr_supported(a), "attn_bwd_hr: unsupported shape L=%d D=%d H=%d r=%d", a.L, a.D, a.H, a.r);
  PEVIT_REQUIRE(ld_dqkv % 8 == 0, "attn_bwd_hr: ld_dqkv=%d must be a multiple of 8", ld_dqkv);
  const HrGeom g = make_geom(a);
  CUtensorMap tq, tk, tv, tdo, to, tqt, tkt, tvt, tdot, tot;
  if (make_tmap_bf16_hm3d(&tq, q, a.L, g.heads, 128) != 0) return -1;
  if (make_tmap_bf16_hm3d(&tk, k, a.L, g.heads, 128) != 0) return -1;
  if (make_tmap_bf16_hm3d(&tv, v, a.L, g.heads, 128) != 0) return -1;
  if (make_tmap_bf16_hm3d(&tqt, q, a.L, g.heads, g.tail16) != 0) return -1;
  if (make_tmap_bf16_hm3d(&tkt, k, a.L, g.heads, g.tail16) != 0) return -1;
  if (make_tmap_bf16_hm3d(&tvt, v, a.L, g.heads, g.tail16) != 0) return -1;
  if (make_tmap_bf16_tok_heads(&tdo, do_tok, a.L, a.NB, a.H, a.D, 128) != 0) return -1;
  if (make_tmap_bf16_tok_heads(&to, o_tok, a.L, a.NB, a.H, a.D, 128) != 0) return -1;
  if (make_tmap_bf16_tok_heads(&tdot, do_tok, a.L, a.NB, a.H, a.D, g.tail16) != 0) return -1;
  if (make_tmap_bf16_tok_heads(&tot, o_tok, a.L, a.NB, a.H, a.D, g.tail16) != 0) return -1;
  // gradient stores: token-major dqkv as [64][3 H column blocks][NB][L], head-major d(delta) as [64][L][2 planes x heads]
  CUtensorMap tg_, tgt, th, tht;
  if (make_tmap_bf16_tok_heads(&tg_, dqkv, a.L, a.NB, 3 * a.H, ld_dqkv, 128) != 0) return -1;
  if (make_tmap_bf16_tok_heads(&tgt, dqkv, a.L, a.NB, 3 * a.H, ld_dqkv, g.tail16) != 0) return -1;
  th = tg_; tht = tgt;
  if (ddelta != nullptr) {
    if (make_tmap_bf16_hm3d(&th, ddelta, a.L, 2 * g.heads, 128) != 0) return -1;
    if (make_tmap_bf16_hm3d(&tht, ddelta, a.L, 2 * g.heads, g.tail16) != 0) return -1;
  }
  const int smem_1box = 4 * g.tensor_bytes + 5 * TILE_BYTES + HB_VEC_BYTES + HB2_MAX_UNITS * static_cast<int>(sizeof(HrUnit)) + 256 + 1024;
  const int nbox = smem_1box + TILE_BYTES <= 227 * 1024 ? 2 : 1;
  const int smem_bytes = smem_1box + (nbox - 1) * TILE_BYTES;
  TraceHost trace;
  // measured (ncu, L = 197, N = 512): with the prefetch the kernel reads 1.27 GB from DRAM, without it 0.78 GB (= the algorithmic
  // 0.775 GB) at the SAME duration (545 us): a head lasts ~23 k cycles, the prefetched lines are evicted before the loads use them
  static const int l2_prefetch = getenv("PEVIT_ATTN_BWD_PF") ? atoi(getenv("PEVIT_ATTN_BWD_PF")) : 0;
  BwdHrParams p{g, ld_dqkv, nbox, l2_prefetch, lse, dqkv, ddelta, trace.begin(HB_THREADS / 32)};
  static const bool use_v1 = getenv("PEVIT_ATTN_BWD_V1") != nullptr;  // diagnostics: the serialised first version
  const int grid = g.heads < sm_count() ? g.heads : sm_count();
  static bool configured[64] = {};
  int dev = 0;
  PEVIT_CHECK_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    PEVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_hr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    PEVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_hr2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[dev & 63] = true;
  }
  ProfScope prof(s, PC_ATTN_BWD);
  if (use_v1)
    PEVIT_CHECK_CUDA(launch_kernel(attn_bwd_hr_kernel, dim3(grid), dim3(HB_THREADS), smem_bytes, s, 1,
                                   tq, tk, tv, tdo, to, tqt, tkt, tvt, tdot, tot, p));
  else
    PEVIT_CHECK_CUDA(launch_kernel(attn_bwd_hr2_kernel, dim3(grid), dim3(HB_THREADS), smem_bytes, s, 1,
                                   tq, tk, tv, tdo, to, tqt, tkt, tvt, tdot, tot, tg_, tgt, th, tht, p));
  PEVIT_CHECK_LAUNCH();
  trace.end(s, "bwd");
  return 0;
}

}  // namespace pevit
