// C[M,N] = A[M,K] * B[N,K]^T on the 5th-gen tensor cores (tcgen05.mma, fp32 accumulators in
// TMEM), bf16 operands staged by TMA into 128B-swizzled shared memory.  Persistent,
// warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 = epilogue
// (TMEM -> registers -> fused epilogue -> global).  Two TMEM accumulator stages let the
// epilogue of tile i overlap the mainloop of tile i+1.
//
// Replaces on the hot path (reference evaluation/model.py): the in-projection `linear`
// (:305) + head split/scale (:729-740,786-787) + the X*P half of adapter_forward (:563-584)
// [EPI_QKV]; out-proj / c_proj `linear` + residual add (:816, :973-974) [EPI_F32];
// c_fc + QuickGELU (:958-962, :165) [EPI_ACT]; and every dgrad GEMM autograd would run.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace pevit {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_THREADS = 320;  // TMA warp, MMA warp, 2 x 4 epilogue warps

// Boxes per epilogue warpgroup (RING): 3 = auxiliary operands are TMA-loaded two boxes ahead (short-K GEMMs whose
// epilogue is the critical path: out-proj + residual, c_proj dgrad x act'(z), the in-place delta); 2 = one box
// ahead / store ring only, which leaves 32 KB more shared memory for mainloop stages (measured: every stage is worth
// ~5 % on the K = 3072 GEMMs).  The host picks per launch (ring_for).
constexpr bool epi_has_ring3(int epi) { return epi == EPI_F32 || epi == EPI_BF16 || epi == EPI_DACT; }

template <int BN, bool CTA2, bool TS, int RING>
struct TileCfg {
  // CTA2: two CTAs (one TPC) share a 256 x BN tile; each stages its own 128 A rows and HALF of B's rows
  static constexpr int B_ROWS = CTA2 ? BN / 2 : BN;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = TS ? 2 * RING * 16384 : 0;  // staged epilogue: 2 warpgroups x RING [128 rows][128 B] boxes
  static constexpr int STAGES_FIT = (227 * 1024 - STAGING_BYTES - 2048) / STAGE_BYTES;
#ifndef PEVIT_GEMM_MAX_STAGES
#define PEVIT_GEMM_MAX_STAGES 8
#endif
  static constexpr int STAGES = STAGES_FIT > PEVIT_GEMM_MAX_STAGES ? PEVIT_GEMM_MAX_STAGES : STAGES_FIT;
  static constexpr int TMEM_COLS = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));  // two accumulator stages, power of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 256 + 1024;  // + barriers + align slack
};

// The activation kind rides in the upper bits of the EPI template argument so that each kernel
// instantiation contains exactly one activation (a runtime switch made ptxas predicate all three).
constexpr int epi_base(int epi) { return epi & 0xff; }
constexpr int epi_act(int epi) { return epi >> 8; }
constexpr int epi_with_act(int epi, int act) { return epi | (act << 8); }

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid(x) = 0.5 tanh(x/2) + 0.5: one MUFU op instead of ex2 + rcp
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

// activations of the path: QuickGELU (model.py:163-165), ReLU (adapter_model.py:305),
// gelu_new (compacter_model.py:374 -> transformers NewGELUActivation)
template <int KIND>
__device__ __forceinline__ float act_fwd(float z) {
  if constexpr (KIND == ACT_QUICKGELU) {
    return z * sigmoid_fast(1.702f * z);
  } else if constexpr (KIND == ACT_RELU) {
    return fmaxf(z, 0.f);
  } else {
    const float u = 0.7978845608028654f * (z + 0.044715f * z * z * z);
    return 0.5f * z * (1.f + tanh_fast(u));
  }
}
template <int KIND>
__device__ __forceinline__ float act_bwd(float z) {
  if constexpr (KIND == ACT_QUICKGELU) {
    const float s = sigmoid_fast(1.702f * z);
    return s * (1.f + 1.702f * z * (1.f - s));
  } else if constexpr (KIND == ACT_RELU) {
    return z > 0.f ? 1.f : 0.f;
  } else {
    const float u = 0.7978845608028654f * (z + 0.044715f * z * z * z);
    const float t = tanh_fast(u);
    return 0.5f * (1.f + t) + 0.5f * z * (1.f - t * t) * 0.7978845608028654f * (1.f + 3.f * 0.044715f * z * z);
  }
}

// Epilogue for one 32-column chunk of one accumulator row.  v[] holds raw fp32 bits.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmEpilogue& ep, int m, int c0,
                                               uint32_t (&v)[32], int N) {
  float a[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) a[j] = __uint_as_float(v[j]);
  const bool full = (c0 + 32 <= N);

  if (ep.bias != nullptr && !(epi_base(EPI) == EPI_QKV && c0 >= 3 * ep.D)) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + c0 + j));
        a[j] += b.x; a[j + 1] += b.y; a[j + 2] += b.z; a[j + 3] += b.w;
      }
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (c0 + j < N) a[j] += __ldg(ep.bias + c0 + j);
    }
  }

  if constexpr (epi_base(EPI) == EPI_F32) {
    const size_t off = static_cast<size_t>(m) * ep.ld_out + c0;
    if (full) {
      if (ep.resid != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 r = *reinterpret_cast<const float4*>(ep.resid + off + j);
          a[j] += r.x; a[j + 1] += r.y; a[j + 2] += r.z; a[j + 3] += r.w;
        }
      }
      if (ep.resid2 != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 r = *reinterpret_cast<const float4*>(ep.resid2 + off + j);
          a[j] += r.x; a[j + 1] += r.y; a[j + 2] += r.z; a[j + 3] += r.w;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(ep.out_f32 + off + j) = make_float4(a[j], a[j + 1], a[j + 2], a[j + 3]);
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (c0 + j < N)
        ep.out_f32[off + j] = a[j] + (ep.resid != nullptr ? ep.resid[off + j] : 0.f) +
                              (ep.resid2 != nullptr ? ep.resid2[off + j] : 0.f);
    }
  } else if constexpr (epi_base(EPI) == EPI_BF16) {
    const size_t off = static_cast<size_t>(m) * ep.ld_out + c0;
    if (ep.resid_bf16 != nullptr) {  // in-place low-rank delta: q' = q + (alpha T Q^T + b)
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const uint4 rr = *reinterpret_cast<const uint4*>(ep.resid_bf16 + off + j);
          const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = unpack_bf16(rw[t]);
            a[j + 2 * t] += f.x;
            a[j + 2 * t + 1] += f.y;
          }
        }
      } else {
        _Pragma("unroll") for (int j = 0; j < 32; ++j) if (c0 + j < N) a[j] += __bfloat162float(ep.resid_bf16[off + j]);
      }
    }
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 p = make_uint4(pack_bf16(a[j], a[j + 1]), pack_bf16(a[j + 2], a[j + 3]),
                             pack_bf16(a[j + 4], a[j + 5]), pack_bf16(a[j + 6], a[j + 7]));
        *reinterpret_cast<uint4*>(ep.out_bf16 + off + j) = p;
      }
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (c0 + j < N) ep.out_bf16[off + j] = __float2bfloat16(a[j]);
    }
  } else if constexpr (epi_base(EPI) == EPI_ACT) {
    // z = acc + bias (kept for backward in out2), out = act(z)
    const size_t off = static_cast<size_t>(m) * ep.ld_out + c0;
    float h[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) h[j] = act_fwd<epi_act(EPI)>(a[j]);
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        *reinterpret_cast<uint4*>(ep.out_bf16 + off + j) =
            make_uint4(pack_bf16(h[j], h[j + 1]), pack_bf16(h[j + 2], h[j + 3]),
                       pack_bf16(h[j + 4], h[j + 5]), pack_bf16(h[j + 6], h[j + 7]));
        if (ep.out2_bf16 != nullptr)
          *reinterpret_cast<uint4*>(ep.out2_bf16 + off + j) =
              make_uint4(pack_bf16(a[j], a[j + 1]), pack_bf16(a[j + 2], a[j + 3]),
                         pack_bf16(a[j + 4], a[j + 5]), pack_bf16(a[j + 6], a[j + 7]));
      }
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (c0 + j < N) {
        ep.out_bf16[off + j] = __float2bfloat16(h[j]);
        if (ep.out2_bf16 != nullptr) ep.out2_bf16[off + j] = __float2bfloat16(a[j]);
      }
    }
  } else if constexpr (epi_base(EPI) == EPI_DACT) {
    // dz = acc * act'(z)
    const size_t off = static_cast<size_t>(m) * ep.ld_out + c0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 zz = *reinterpret_cast<const uint4*>(ep.aux_bf16 + off + j);
        uint32_t zw[4] = {zz.x, zz.y, zz.z, zz.w};
        uint32_t o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float2 z = unpack_bf16(zw[t]);
          o[t] = pack_bf16(a[j + 2 * t] * act_bwd<epi_act(EPI)>(z.x), a[j + 2 * t + 1] * act_bwd<epi_act(EPI)>(z.y));
        }
        *reinterpret_cast<uint4*>(ep.out_bf16 + off + j) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (c0 + j < N) {
        float z = __bfloat162float(ep.aux_bf16[off + j]);
        ep.out_bf16[off + j] = __float2bfloat16(a[j] * act_bwd<epi_act(EPI)>(z));
      }
    }
  } else if constexpr (epi_base(EPI) == EPI_QKV) {
    // rows are tokens in LND order (m = l*NB + n).  Columns [0,3D): q|k|v -> head-major bf16
    // tiles [which][n*H+h][l][64] with q pre-scaled by 1/sqrt(64) (exact: power of two);
    // columns [3D, 3D+r2): low-rank activations T = X*P, stored bf16 row-major [M][r2] (A operand of the delta GEMM).
    const int l = m / ep.NB, n = m - l * ep.NB;
    const int threeD = 3 * ep.D;
    if (c0 < threeD) {
      const int which = c0 / ep.D;
      const int within = c0 - which * ep.D;
      const int h = within >> 6, d0 = within & 63;
      const float sc = which == 0 ? 0.125f : 1.f;
      bf16* dst = ep.qkv_hm +
                  ((static_cast<size_t>(which) * ep.NB * ep.H + static_cast<size_t>(n) * ep.H + h) * ep.L + l) * 64 + d0;
#pragma unroll
      for (int j = 0; j < 32; j += 8)
        *reinterpret_cast<uint4*>(dst + j) =
            make_uint4(pack_bf16(a[j] * sc, a[j + 1] * sc), pack_bf16(a[j + 2] * sc, a[j + 3] * sc),
                       pack_bf16(a[j + 4] * sc, a[j + 5] * sc), pack_bf16(a[j + 6] * sc, a[j + 7] * sc));
    } else {
      const int t0 = c0 - threeD;
      bf16* dst = ep.t_out + static_cast<size_t>(m) * ep.r2 + t0;
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (t0 + j < ep.r2) dst[j] = __float2bfloat16(a[j]);
    }
  }
}

// Staged ("box") epilogue.  A box is [128 rows][128 B] of shared memory in the 128B-swizzled layout TMA reads and
// writes: 32 fp32 or 64 bf16 output columns of the tile.  An auxiliary operand of the epilogue (fp32 residual, bf16
// residual of the in-place delta, saved pre-activation z) is TMA-loaded INTO the box, combined in place with the
// accumulator by the thread that owns the row, and the box leaves through a TMA store -- every global access of the
// epilogue is a bulk, coalesced, asynchronous copy.
constexpr int BOX_BYTES = 16384;

// bias of 32 consecutive columns, fetched before the TMEM load is waited for so that its latency overlaps it
struct BiasRegs { float4 b[8]; };
template <int EPI>
__device__ __forceinline__ void load_bias(const GemmEpilogue& ep, int c0, BiasRegs& br) {
  const bool on = ep.bias != nullptr && !(epi_base(EPI) == EPI_QKV && c0 >= 3 * ep.D);
#pragma unroll
  for (int q = 0; q < 8; ++q)
    br.b[q] = on ? __ldg(reinterpret_cast<const float4*>(ep.bias + c0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// QuickGELU forward of 32 columns: packed bf16 h = act(z) and z = acc + bias
template <int EPI>
__device__ __forceinline__ void act_half_compute(const BiasRegs& br, const uint32_t (&v)[32], uint32_t (&hp)[16],
                                                 uint32_t (&zp)[16]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float z0 = __uint_as_float(v[4 * q]) + br.b[q].x, z1 = __uint_as_float(v[4 * q + 1]) + br.b[q].y;
    const float z2 = __uint_as_float(v[4 * q + 2]) + br.b[q].z, z3 = __uint_as_float(v[4 * q + 3]) + br.b[q].w;
    zp[2 * q] = pack_bf16(z0, z1);
    zp[2 * q + 1] = pack_bf16(z2, z3);
    hp[2 * q] = pack_bf16(act_fwd<epi_act(EPI)>(z0), act_fwd<epi_act(EPI)>(z1));
    hp[2 * q + 1] = pack_bf16(act_fwd<epi_act(EPI)>(z2), act_fwd<epi_act(EPI)>(z3));
  }
}
__device__ __forceinline__ void act_half_store(uint8_t* rowp, uint8_t* rowp2, int r7, int sub, const uint32_t (&hp)[16],
                                               const uint32_t (&zp)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int off = ((sub * 4 + q) ^ r7) << 4;
    *reinterpret_cast<uint4*>(rowp + off) = make_uint4(hp[4 * q], hp[4 * q + 1], hp[4 * q + 2], hp[4 * q + 3]);
    if (rowp2 != nullptr)
      *reinterpret_cast<uint4*>(rowp2 + off) = make_uint4(zp[4 * q], zp[4 * q + 1], zp[4 * q + 2], zp[4 * q + 3]);
  }
}

// 32 accumulator columns of one row -> the row's slice of the box (all epilogues but EPI_ACT).  `rowp` = box +
// row * 128, `r7` = row & 7, `sub` = which 64-byte half of the row (bf16 boxes hold two 32-column halves; fp32
// boxes one: the whole row).
template <int EPI>
__device__ __forceinline__ void box_half(const BiasRegs& br, bool has_aux, uint8_t* rowp, int r7, int sub, float scale,
                                         const uint32_t (&v)[32]) {
  float a[32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    a[4 * q] = __uint_as_float(v[4 * q]) + br.b[q].x;
    a[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + br.b[q].y;
    a[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + br.b[q].z;
    a[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + br.b[q].w;
  }
  if constexpr (epi_base(EPI) == EPI_F32) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4* p = reinterpret_cast<float4*>(rowp + ((q ^ r7) << 4));
      float4 o = make_float4(a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
      if (has_aux) { const float4 r = *p; o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
      *p = o;
    }
  } else {  // EPI_BF16 (+ in-place bf16 residual), EPI_DACT (x act'(z)), EPI_QKV (x scale)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4* p = reinterpret_cast<uint4*>(rowp + (((sub * 4 + q) ^ r7) << 4));
      if (epi_base(EPI) == EPI_DACT || (epi_base(EPI) == EPI_BF16 && has_aux)) {
        const uint4 x = *p;
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = unpack_bf16(w[t]);
          if constexpr (epi_base(EPI) == EPI_DACT) {
            a[8 * q + 2 * t] *= act_bwd<epi_act(EPI)>(f.x);
            a[8 * q + 2 * t + 1] *= act_bwd<epi_act(EPI)>(f.y);
          } else {
            a[8 * q + 2 * t] += f.x;
            a[8 * q + 2 * t + 1] += f.y;
          }
        }
      }
      if constexpr (epi_base(EPI) == EPI_QKV) {
#pragma unroll
        for (int t = 0; t < 8; ++t) a[8 * q + t] *= scale;
      }
      *p = make_uint4(pack_bf16(a[8 * q], a[8 * q + 1]), pack_bf16(a[8 * q + 2], a[8 * q + 3]),
                      pack_bf16(a[8 * q + 4], a[8 * q + 5]), pack_bf16(a[8 * q + 6], a[8 * q + 7]));
    }
  }
}

template <int BN, int EPI, bool TS, bool CTA2, int RING>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_c2,
               const __grid_constant__ CUtensorMap tmap_aux, int M, int N, int K, GemmEpilogue ep) {
  using Cfg = TileCfg<BN, CTA2, TS, RING>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int TILE_M = CTA2 ? 2 * BM : BM;  // rows of C covered by one (pair of) CTA(s) per tile
  // staged epilogue geometry (see box_half): columns per box, boxes per tile
  constexpr bool kF32 = (epi_base(EPI) == EPI_F32);
  constexpr bool kTwo = (epi_base(EPI) == EPI_ACT);  // h and z leave together, one box each
  constexpr bool kQKV = (epi_base(EPI) == EPI_QKV);
  constexpr int BOXCOLS = kF32 ? 32 : 64;
  constexpr int NBOXES = BN / BOXCOLS;
  static_assert(!TS || NBOXES >= 1, "staged epilogue needs at least one box per tile");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;  // 1024-aligned: STAGE_BYTES is a multiple of 1024
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* aux_bar = tempty_bar + 2;  // [2 warpgroups][3 boxes]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = (N + BN - 1) / BN;
  const int tiles_m = (M + TILE_M - 1) / TILE_M;
  const int tiles_per_batch = tiles_m * tiles_n;
  const int num_tiles = tiles_per_batch * ep.batch;
  const int num_kb = (K + BK - 1) / BK;
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;       // 0 = leader (issues the MMAs)
  const int first_tile = CTA2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = CTA2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  // Batched launch (ep.batch > 1): several products that share shapes run as one grid; tile index = batch-major.
  // Operands / outputs of batch b sit at fixed ROW offsets of the same 2-D tensors (plus a pointer offset for the
  // direct-store epilogue).
  auto tile_batch = [&](int tile) { return tile / tiles_per_batch; };
  auto tile_m0 = [&](int tile) { return ((tile % tiles_per_batch) / tiles_n) * TILE_M + static_cast<int>(cta_rank) * BM; };
  auto tile_n0 = [&](int tile) { return ((tile % tiles_per_batch) % tiles_n) * BN; };
  // auxiliary epilogue operand that is TMA-loaded into the boxes (staged epilogue only)
  const bool has_aux = TS && (kF32 ? ep.resid != nullptr
                                   : (epi_base(EPI) == EPI_BF16 ? ep.resid_bf16 != nullptr : epi_base(EPI) == EPI_DACT));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if constexpr (TS) {
      tma_prefetch_desc(&tmap_c);
      if (kTwo || kQKV) tma_prefetch_desc(&tmap_c2);
      if (has_aux) tma_prefetch_desc(&tmap_aux);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], CTA2 ? 16 : 8);  // one arrive per epilogue warp (of both CTAs)
    }
    for (int s = 0; s < 6; ++s) mbar_init(&aux_bar[s], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (CTA2) { tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();  // the peer's barriers exist before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // only after the TMEM allocation: a co-resident successor must not take the columns first
  pdl_wait();               // predecessor grid complete and visible; nothing above touches global memory

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int bidx = tile_batch(tile);
        const int m0 = tile_m0(tile) + bidx * ep.a_batch_rows;
        const int n0 = tile_n0(tile) + static_cast<int>(cta_rank) * Cfg::B_ROWS * (CTA2 ? 1 : 0) + bidx * ep.b_batch_rows;
        if (has_aux) {
          // pull this tile's auxiliary boxes into L2 a whole mainloop ahead of the epilogue that consumes them
          const int nt = tile_n0(tile);
#pragma unroll 1
          for (int j = 0; j < NBOXES; ++j)
            if (nt + j * BOXCOLS < N) tma_prefetch_l2_2d(&tmap_aux, nt + j * BOXCOLS, tile_m0(tile) + bidx * ep.c_batch_rows);
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          // A (activations, tens of MB, read once per N tile) mostly misses L2; the ring buffers ~1 us of mainloop,
          // less than an HBM round trip under load -- pull the A box of a later k-block into L2 now (B, the frozen
          // weights, stays L2-resident on its own)
          if (ep.a_prefetch > 0 && kb + ep.a_prefetch < num_kb) tma_prefetch_l2_2d(&tmap_a, (kb + ep.a_prefetch) * BK, m0);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          if constexpr (CTA2) {
            // both CTAs' loads credit the leader's barrier, which is armed for the bytes of the whole pair
            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], kb * BK, m0);
            tma_load_2d_pair(sa + Cfg::A_BYTES, &tmap_b, &full_bar[stage], kb * BK, n0);
          } else if (ep.debug & 1) {
            mbar_arrive(&full_bar[stage]);  // diagnostics: pipeline without operand traffic
          } else {
            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m0);
            tma_load_2d(sa + Cfg::A_BYTES, &tmap_b, &full_bar[stage], kb * BK, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only when paired)
    constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = first_tile; tile < num_tiles && cta_rank == 0; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t da = umma_desc_kmajor_sw128(sa);
          const uint64_t db = umma_desc_kmajor_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16 && !(ep.debug & 2); ++k) {  // +32 B per 16-element K step (encoded >> 4)
            if constexpr (CTA2) umma_bf16_ss_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            else umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          if constexpr (CTA2) {
            umma_commit_pair(&empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit_pair(&tfull_bar[acc]);
          } else if (ep.debug & 2) {
            mbar_arrive(&empty_bar[stage]);  // diagnostics: operand traffic without tensor-core work
            if (kb == num_kb - 1) mbar_arrive(&tfull_bar[acc]);
          } else {
            umma_commit(&empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: two warpgroups (warps 2-5, 6-9)
    // Both warpgroups cover all 128 accumulator rows (TMEM lane quadrant = warp % 4).
    const int quad = warp & 3;          // TMEM lane quadrant this warp may read
    const int wg = (warp - 2) >> 2;     // 0 or 1
    if constexpr (!TS) {
      // direct stores: the warpgroups take alternate 32-column chunks of every tile
      int it = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const int m_base = tile_m0(tile);
        const int m = m_base + quad * 32 + lane;
        const int n0 = tile_n0(tile);
        GemmEpilogue epb = ep;  // batch b writes at a fixed element offset of the same output
        if (ep.batch > 1 && epb.out_bf16 != nullptr) epb.out_bf16 += static_cast<size_t>(tile_batch(tile)) * ep.c_batch_elems;
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
#pragma unroll 1
        for (int c = wg * 32; c < BN; c += 64) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + c, v);
          tmem_ld_wait();
          if (m < M && n0 + c < N) epilogue_chunk<EPI>(epb, m, n0 + c, v, N);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CTA2) mbar_arrive_cluster_relaxed(&tempty_bar[acc], 0);  // the leader's MMA warp owns the accumulators
          else mbar_arrive(&tempty_bar[acc]);
        }
      }
    } else {
      // Staged epilogue.  Boxes of the CTA's tile sequence are numbered b = it * NBOXES + j; warpgroup wg owns the
      // boxes with b % 2 == wg and a private ring of RING box buffers:
      //   [aux TMA load -> box] -> accumulator (+bias) combined in place -> fence -> barrier -> TMA store
      // RING = 3 (aux-capable epilogues): the aux load of box g+2 is issued right after the store of box g, once
      // the store of box g-1 (same buffer) has drained.  RING = 2: the elected thread drains the previous store
      // before the barrier, so the other buffer is free for the next box.
      uint8_t* my_boxes = staging + wg * RING * BOX_BYTES;
      uint64_t* my_aux = aux_bar + wg * 3;
      const int row = quad * 32 + lane, r7 = row & 7;
      // One thread of the warpgroup's first warp owns the box traffic (bulk groups are per thread).  elect.sync picks the
      // same lane for the same mask every time, and tells ptxas that the guarded code runs on ONE lane (plain R2URs for
      // the TMA operands instead of a waterfall loop per instruction).
      const bool first_warp = warp == 2 + 4 * wg;
      auto elected = [&]() { return first_warp && elect_one(); };
      const int bar_id = 1 + wg;
      auto first_j = [&](int it_) { return (wg ^ ((it_ * NBOXES) & 1)) & 1; };
      // cursor over this warpgroup's boxes, ahead of the consumer (aux prefetch)
      int la_it = 0, la_j = first_j(0);
      auto la_seek = [&]() -> bool {
        for (;;) {
          const int tile_ = first_tile + la_it * tile_step;
          if (tile_ >= num_tiles) return false;
          if (la_j < NBOXES && tile_n0(tile_) + la_j * BOXCOLS < N) return true;
          ++la_it;
          la_j = first_j(la_it);
        }
      };
      auto la_issue = [&](int slot) {  // elected only
        if (!la_seek()) return;
        const int tile_ = first_tile + la_it * tile_step;
        const int m0_ = tile_m0(tile_) + tile_batch(tile_) * ep.c_batch_rows;
        const int c0_ = tile_n0(tile_) + la_j * BOXCOLS;
        mbar_expect_tx(&my_aux[slot], BOX_BYTES);
        tma_load_2d(my_boxes + slot * BOX_BYTES, &tmap_aux, &my_aux[slot], c0_, m0_);
        la_j += 2;
      };
      if (has_aux && elected()) {
#pragma unroll
        for (int q = 0; q < RING - 1; ++q) la_issue(q);
      }
      int slot = 0;            // (boxes processed by this warpgroup) % RING
      uint32_t aux_phase = 0;  // (boxes processed / RING) & 1
      int it = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const int m0 = tile_m0(tile) + tile_batch(tile) * ep.c_batch_rows;  // output / aux rows (QKV: batch == 1)
        const int n0 = tile_n0(tile);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
#pragma unroll 1
        for (int j = first_j(it); j < NBOXES; j += 2) {
          const int c0 = n0 + j * BOXCOLS;
          if (c0 >= N) break;
          uint8_t* box = my_boxes + (kTwo ? 0 : slot) * BOX_BYTES;
          uint8_t* rowp = box + row * 128;
          if constexpr (kTwo) {
            uint8_t* rowp2 = ep.out2_bf16 != nullptr ? rowp + BOX_BYTES : nullptr;
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              uint32_t v[32], hp[16], zp[16];
              BiasRegs br;
              tmem_ld_32x32(t_row + j * 64 + sub * 32, v);
              load_bias<EPI>(ep, c0 + sub * 32, br);
              tmem_ld_wait();
              act_half_compute<EPI>(br, v, hp, zp);
              if (sub == 0) {  // both buffers are about to be rewritten: the previous h / z stores must have drained
                if (elected()) tma_store_wait_read<0>();
                named_bar_sync(bar_id, 128);
              }
              act_half_store(rowp, rowp2, r7, sub, hp, zp);
            }
          } else {
            float scale = 1.f;
            if constexpr (kQKV) scale = c0 < ep.D ? 0.125f : 1.f;
#pragma unroll
            for (int sub = 0; sub < (kF32 ? 1 : 2); ++sub) {
              uint32_t v[32];
              BiasRegs br;
              tmem_ld_32x32(t_row + j * BOXCOLS + sub * 32, v);
              load_bias<EPI>(ep, c0 + sub * 32, br);
              if (sub == 0 && has_aux) mbar_wait(&my_aux[slot], aux_phase);
              tmem_ld_wait();
              if (!(ep.debug & 4)) box_half<EPI>(br, has_aux, rowp, r7, sub, scale, v);
            }
          }
          fence_proxy_async_smem();
          if constexpr (!kTwo) {
            // the buffer of the NEXT box must be free once the barrier below is passed
            if (elected()) tma_store_wait_read<RING - 2>();
          }
          named_bar_sync(bar_id, 128);
          if (elected()) {
            if constexpr (kQKV) {
              if (c0 < 3 * ep.D) {
                const int which = c0 / ep.D, h = (c0 - which * ep.D) >> 6;
                tma_store_5d(&tmap_c, box, 0, m0 / ep.NB, h, m0 % ep.NB, which);  // [64 d][l][h][n][which]
              } else {
                tma_store_2d(&tmap_c2, box, c0 - 3 * ep.D, m0);                    // T columns
              }
            } else {
              tma_store_2d(&tmap_c, box, c0, m0);
              if constexpr (kTwo) { if (ep.out2_bf16 != nullptr) tma_store_2d(&tmap_c2, box + BOX_BYTES, c0, m0); }
            }
            tma_store_commit();
            if (has_aux) {
              tma_store_wait_read<1>();  // store of box g-1 drained: its buffer takes the aux operand of box g+RING-1
              la_issue(slot == 0 ? RING - 1 : slot - 1);
            }
          }
          if (++slot == RING) { slot = 0; aux_phase ^= 1; }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CTA2) mbar_arrive_cluster_relaxed(&tempty_bar[acc], 0);  // the leader's MMA warp owns the accumulators
          else mbar_arrive(&tempty_bar[acc]);
        }
      }
      if (elected()) tma_store_wait_all<0>();  // smem must outlive the last stores
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();  // no CTA may exit (or free TMEM) while its peer still signals it
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int EPI, bool TS, bool CTA2, int RING>
int launch_impl(cudaStream_t stream, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                const CUtensorMap& tc2, const CUtensorMap& taux, int M, int N, int K, const GemmEpilogue& ep) {
  using Cfg = TileCfg<BN, CTA2, TS, RING>;
  static_assert(Cfg::STAGES >= 2, "pipeline too shallow");
  static bool configured[64] = {};  // per instantiation and device (the attribute is per context)
  auto kern = gemm_tn_kernel<BN, EPI, TS, CTA2, RING>;
  int dev = 0;
  PEVIT_CHECK_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    PEVIT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured[dev & 63] = true;
  }
  const int tile_m = CTA2 ? 2 * BM : BM;
  const int tiles = ((M + tile_m - 1) / tile_m) * ((N + BN - 1) / BN) * ep.batch;
  ProfScope prof(stream, PC_GEMM_OTHER);
  int grid;
  if constexpr (CTA2) {
    const int pairs = sm_count() / 2;
    grid = 2 * (tiles < pairs ? tiles : pairs);
  } else {
    grid = tiles < sm_count() ? tiles : sm_count();
  }
  PEVIT_CHECK_CUDA(launch_kernel(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, CTA2 ? 2 : 1, ta, tb, tc,
                                 tc2, taux, M, N, K, ep));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

// CTA pairs are used for the wide tiles of the big GEMMs; narrow tiles (skinny N) stay single-CTA.
thread_local bool g_use_pair = false;

struct Tmaps { CUtensorMap a, b, c, c2, aux; };

thread_local int g_ring = 2;

template <int BN, int EPI, bool TS_REQ>
int launch(cudaStream_t stream, const Tmaps& t, int M, int N, int K, const GemmEpilogue& ep) {
  // a tile narrower than one box (32 fp32 / 64 bf16 columns) cannot use the staged epilogue
  constexpr bool TS = TS_REQ && BN >= (epi_base(EPI) == EPI_F32 ? 32 : 64);
  if constexpr (TS && epi_has_ring3(epi_base(EPI))) {
    if (g_ring == 3) {
      if constexpr (BN >= 128) {
        if (g_use_pair) return launch_impl<BN, EPI, TS, true, 3>(stream, t.a, t.b, t.c, t.c2, t.aux, M, N, K, ep);
      }
      return launch_impl<BN, EPI, TS, false, 3>(stream, t.a, t.b, t.c, t.c2, t.aux, M, N, K, ep);
    }
  }
  if constexpr (BN >= 128) {
    if (g_use_pair) return launch_impl<BN, EPI, TS, true, 2>(stream, t.a, t.b, t.c, t.c2, t.aux, M, N, K, ep);
  }
  return launch_impl<BN, EPI, TS, false, 2>(stream, t.a, t.b, t.c, t.c2, t.aux, M, N, K, ep);
}

template <int BN>
int dispatch_epi(int epi, bool ts, cudaStream_t s, const Tmaps& t, int M, int N, int K, const GemmEpilogue& ep) {
  switch (epi) {
    case EPI_F32: return ts ? launch<BN, EPI_F32, true>(s, t, M, N, K, ep) : launch<BN, EPI_F32, false>(s, t, M, N, K, ep);
    case EPI_BF16: return ts ? launch<BN, EPI_BF16, true>(s, t, M, N, K, ep) : launch<BN, EPI_BF16, false>(s, t, M, N, K, ep);
    case EPI_ACT:
      if (ep.act == ACT_QUICKGELU)
        return ts ? launch<BN, epi_with_act(EPI_ACT, ACT_QUICKGELU), true>(s, t, M, N, K, ep)
                  : launch<BN, epi_with_act(EPI_ACT, ACT_QUICKGELU), false>(s, t, M, N, K, ep);
      if (ep.act == ACT_RELU) return launch<BN, epi_with_act(EPI_ACT, ACT_RELU), false>(s, t, M, N, K, ep);
      return launch<BN, epi_with_act(EPI_ACT, ACT_GELU_NEW), false>(s, t, M, N, K, ep);
    case EPI_DACT:
      if (ep.act == ACT_QUICKGELU)
        return ts ? launch<BN, epi_with_act(EPI_DACT, ACT_QUICKGELU), true>(s, t, M, N, K, ep)
                  : launch<BN, epi_with_act(EPI_DACT, ACT_QUICKGELU), false>(s, t, M, N, K, ep);
      if (ep.act == ACT_RELU) return launch<BN, epi_with_act(EPI_DACT, ACT_RELU), false>(s, t, M, N, K, ep);
      return launch<BN, epi_with_act(EPI_DACT, ACT_GELU_NEW), false>(s, t, M, N, K, ep);
    case EPI_QKV: return ts ? launch<BN, EPI_QKV, true>(s, t, M, N, K, ep) : launch<BN, EPI_QKV, false>(s, t, M, N, K, ep);
  }
  set_error("gemm_tn: unknown epilogue %d", epi);
  return -1;
}

// Tile choice.  Tiles are operand-fetch (L2 -> SM) bound, so the cost of a schedule is (number of waves) x (bytes
// one CTA stages per k-block): BM + BN rows single-CTA, BM + BN/2 rows when a CTA pair shares the B tile.
int pick_tile(int M, int N, int forced, bool allow_pair, bool* pair) {
  *pair = false;
  if (forced < 0) forced = -forced;  // negative: same tile width, direct-store epilogue (cross-check)
  const bool forced_pair = forced >= 1000;
  if (forced_pair) forced -= 1000;
  if (forced == 32 || forced == 64 || forced == 128 || forced == 192 || forced == 256) {
    *pair = forced_pair && forced >= 128;
    return forced;
  }
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  const int sms = sm_count();
  int best = 256;
  double best_cost = 1e30;
  for (int use_pair = (allow_pair && M > BM) ? 1 : 0; use_pair >= 0; --use_pair) {
    for (int bn : {256, 192, 128, 64}) {
      if (use_pair && bn < 128) continue;
      const int tile_m = use_pair ? 2 * BM : BM;
      const int units = use_pair ? sms / 2 : sms;
      const int tiles = ((M + tile_m - 1) / tile_m) * ((N + bn - 1) / bn);
      const int waves = (tiles + units - 1) / units;
      const double cost = static_cast<double>(waves) * (BM + (use_pair ? bn / 2 : bn));
      if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; *pair = use_pair != 0; }
    }
  }
  return best;
}

}  // namespace

int gemm_tn(cudaStream_t stream, const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K, int epi,
            const GemmEpilogue& ep_in, int force_bn) {
  static const int a_prefetch = getenv("PEVIT_GEMM_APF") ? atoi(getenv("PEVIT_GEMM_APF")) : 0;
  GemmEpilogue ep = ep_in;
  ep.a_prefetch = a_prefetch;
  PEVIT_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_tn: empty problem %d x %d x %d", M, N, K);
  PEVIT_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0,
                "gemm_tn: K, lda, ldb must be multiples of 8 (16-byte TMA strides): K=%d lda=%d ldb=%d", K, lda, ldb);
  PEVIT_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
                "gemm_tn: operands must be 16-byte aligned");
  if (epi == EPI_QKV)
    PEVIT_REQUIRE(ep.D % 64 == 0 && N == 3 * ep.D + ep.r2 && M == ep.L * ep.NB && ep.H * 64 == ep.D,
                  "gemm_tn[qkv]: inconsistent shape M=%d N=%d D=%d r2=%d L=%d NB=%d H=%d", M, N, ep.D, ep.r2, ep.L,
                  ep.NB, ep.H);
  else
    PEVIT_REQUIRE(ep.ld_out % 8 == 0, "gemm_tn: ld_out must be a multiple of 8");
  bool pair = false;
  static const bool pair_disabled = getenv("PEVIT_GEMM_NO_PAIR") != nullptr;
  const int bn = pick_tile(M, N, force_bn, !pair_disabled, &pair);
  g_use_pair = pair;
  PEVIT_REQUIRE(ep.batch >= 1 && (ep.batch == 1 || epi == EPI_BF16),
                "gemm_tn: batched launches are supported for the bf16 epilogue only (batch=%d, epilogue %d)", ep.batch, epi);
  const int nbm1 = ep.batch - 1;
  Tmaps t;
  if (make_tmap_bf16_2d(&t.a, A, static_cast<uint64_t>(nbm1) * ep.a_batch_rows + M, K, lda, BM, BK) != 0) return -1;
  if (make_tmap_bf16_2d(&t.b, B, static_cast<uint64_t>(nbm1) * ep.b_batch_rows + N, K, ldb, pair ? bn / 2 : bn, BK) != 0) return -1;
  t.c = t.a;
  t.c2 = t.a;
  t.aux = t.a;
  auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  bool ts = false;
  if (epi == EPI_QKV) {
    // staged head-major scatter: a 128-row tile must be 128 images of ONE token index, and the low-rank columns
    // (if any) exactly one 64-column box
    ts = force_bn >= 0 && ep.NB % 128 == 0 && (ep.r2 == 0 || ep.r2 == 64) && aligned16(ep.qkv_hm) &&
         (ep.r2 == 0 || aligned16(ep.t_out));
    if (ts) {
      if (make_tmap_qkv_hm_5d(&t.c, ep.qkv_hm, ep.L, ep.NB, ep.H, BM) != 0) return -1;
      if (ep.r2 != 0 && make_tmap_out_2d(&t.c2, ep.t_out, M, ep.r2, ep.r2, BM, 2) != 0) return -1;
    }
  } else {
    // Staged TMA epilogue for row-major outputs with 16-byte aligned rows; direct stores otherwise.
    const void* out = epi == EPI_F32 ? static_cast<const void*>(ep.out_f32) : static_cast<const void*>(ep.out_bf16);
    const void* aux = epi == EPI_F32 ? static_cast<const void*>(ep.resid)
                                     : (epi == EPI_BF16 ? static_cast<const void*>(ep.resid_bf16)
                                                        : (epi == EPI_DACT ? static_cast<const void*>(ep.aux_bf16) : nullptr));
    const int elt = epi == EPI_F32 ? 4 : 2;
    const bool act_plain = !((epi == EPI_ACT || epi == EPI_DACT) && ep.act != ACT_QUICKGELU);  // bottleneck acts: direct
    ts = act_plain && N % 64 == 0 && out != nullptr && aligned16(out) && aligned16(aux) && ep.resid2 == nullptr &&
         (static_cast<size_t>(ep.ld_out) * elt) % 16 == 0 && aligned16(ep.out2_bf16) && force_bn >= 0;
    if (ep.batch > 1) {
      // staged: batches are ROW blocks of one output tensor, so a tile must not straddle two of them;
      // direct: batches are element offsets of the same rows
      const int tile_m = pair ? 2 * BM : BM;
      if (ts) PEVIT_REQUIRE(ep.c_batch_elems == 0 && ep.c_batch_rows >= M && M % tile_m == 0,
                            "gemm_tn: batched staged epilogue needs M %% %d == 0 and row-block outputs (M=%d)", tile_m, M);
      else PEVIT_REQUIRE(ep.c_batch_rows == 0, "gemm_tn: batched direct epilogue takes c_batch_elems, not c_batch_rows");
    }
    const uint64_t out_rows = static_cast<uint64_t>(nbm1) * ep.c_batch_rows + M;
    if (ts) {
      if (make_tmap_out_2d(&t.c, out, out_rows, N, ep.ld_out, BM, elt) != 0) return -1;
      if (epi == EPI_ACT && ep.out2_bf16 != nullptr && make_tmap_out_2d(&t.c2, ep.out2_bf16, M, N, ep.ld_out, BM, 2) != 0)
        return -1;
      if (aux != nullptr && make_tmap_out_2d(&t.aux, aux, out_rows, N, ep.ld_out, BM, elt) != 0) return -1;
    }
    // deep aux ring only where the epilogue is the critical path (short K); long-K GEMMs get the stages instead
    static const int ring3_max_k = getenv("PEVIT_GEMM_RING3_MAXK") ? atoi(getenv("PEVIT_GEMM_RING3_MAXK")) : 1024;
    g_ring = (aux != nullptr && K <= ring3_max_k) ? 3 : 2;
  }
  switch (bn) {
    case 32: return dispatch_epi<32>(epi, ts, stream, t, M, N, K, ep);
    case 64: return dispatch_epi<64>(epi, ts, stream, t, M, N, K, ep);
    case 128: return dispatch_epi<128>(epi, ts, stream, t, M, N, K, ep);
    case 192: return dispatch_epi<192>(epi, ts, stream, t, M, N, K, ep);
    default: return dispatch_epi<256>(epi, ts, stream, t, M, N, K, ep);
  }
}

}  // namespace pevit
