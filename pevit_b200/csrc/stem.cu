// ViT stem (reference evaluation/model.py:1034-1042): conv1 with stride == kernel (patch embedding),
// class token, positional embedding, ln_pre and the NLD -> LND permute.  A stride-p convolution is a
// GEMM over non-overlapping patches, so the stem is: im2col + bf16 cast (one pass over the fp32
// images) -> tcgen05 GEMM against the flattened conv weight -> one finalize pass that adds the
// positional embedding / class token, applies ln_pre and writes the (L, N, D) fp32 residual stream.
// Frozen parameters only: no backward (nothing upstream of the first block trains).
#include "common.cuh"
#include "kernels.h"

namespace pevit {
namespace {

// patches[(n*G + gy)*G + gx][c*p*p + i*p + j] = img[n][c][gy*p + i][gx*p + j]   (bf16, K padded with zeros)
// One block per patch.  p % 4 == 0 (every CLIP ViT: 32, 16; not 14): a thread converts 4 consecutive pixels of one
// patch row (16-byte load, 8-byte store); otherwise element by element.
template <bool kVec4>
__global__ void im2col_kernel(const float* __restrict__ img, bf16* __restrict__ patches, int NB, int R, int p, int G,
                              int K, int Kpad) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;  // (n, gy, gx)
  const int n = row / (G * G), g = row - n * G * G;
  const int gy = g / G, gx = g - gy * G;
  const float* src = img + (static_cast<size_t>(n) * 3 * R + gy * p) * R + gx * p;
  bf16* dst = patches + static_cast<size_t>(row) * Kpad;
  const int pp = p * p;
  if constexpr (kVec4) {
    const int p4 = p >> 2;
    for (int q = threadIdx.x; q < (K >> 2); q += blockDim.x) {  // q = (c, i, j/4)
      const int ci = q / p4, j4 = q - ci * p4;
      const int c = ci / p, i = ci - c * p;
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + (static_cast<size_t>(c) * R + i) * R) + j4);
      *reinterpret_cast<uint2*>(dst + 4 * q) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
    for (int k = K + threadIdx.x; k < Kpad; k += blockDim.x) dst[k] = __float2bfloat16(0.f);
  } else {
    for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
      float v = 0.f;
      if (k < K) {
        const int c = k / pp, rem = k - c * pp;
        const int i = rem / p, j = rem - i * p;
        v = __ldg(src + (static_cast<size_t>(c) * R + i) * R + j);
      }
      dst[k] = __float2bfloat16(v);
    }
  }
}

constexpr int FIN_THREADS = 256;
constexpr int FIN_MAXV = 8;

// x[l*NB + n][:] = ln_pre( (l == 0 ? cls : emb[n*G2 + l-1]) + pos[l] )
__global__ void __launch_bounds__(FIN_THREADS)
embed_finalize_kernel(const float* __restrict__ emb, const float* __restrict__ cls, const float* __restrict__ pos,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ x, int NB,
                      int L, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (FIN_THREADS / 32) + warp;  // output row in LND order
  if (row >= L * NB) return;
  const int l = row / NB, n = row - l * NB;
  const int nv = D >> 7;
  const float4* src = reinterpret_cast<const float4*>(l == 0 ? cls : emb + (static_cast<size_t>(n) * (L - 1) + l - 1) * D);
  const float4* ps = reinterpret_cast<const float4*>(pos + static_cast<size_t>(l) * D);
  float4 buf[FIN_MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < FIN_MAXV; ++i)
    if (i < nv) {
      const float4 a = src[lane + 32 * i], b = __ldg(ps + lane + 32 * i);
      buf[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      sum += (buf[i].x + buf[i].y) + (buf[i].z + buf[i].w);
    }
  const float mean = warp_sum(sum) / D;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < FIN_MAXV; ++i)
    if (i < nv) {
      const float a = buf[i].x - mean, b = buf[i].y - mean, c = buf[i].z - mean, d = buf[i].w - mean;
      var += (a * a + b * b) + (c * c + d * d);
    }
  const float rstd = rsqrtf(warp_sum(var) / D + 1e-5f);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  float4* out = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * D);
#pragma unroll
  for (int i = 0; i < FIN_MAXV; ++i)
    if (i < nv) {
      const int c = lane + 32 * i;
      const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
      out[c] = make_float4((buf[i].x - mean) * rstd * g.x + b.x, (buf[i].y - mean) * rstd * g.y + b.y,
                           (buf[i].z - mean) * rstd * g.z + b.z, (buf[i].w - mean) * rstd * g.w + b.w);
    }
}

}  // namespace

size_t patch_embed_workspace_bytes(int NB, int R, int p, int D) {
  const int G = R / p, K = 3 * p * p, Kpad = (K + 7) / 8 * 8;
  const size_t rows = static_cast<size_t>(NB) * G * G;
  return ((rows * Kpad * sizeof(bf16) + 255) & ~size_t(255)) + rows * D * sizeof(float) + 256;
}

int patch_embed(cudaStream_t s, const float* img, const bf16* w_patch, const float* cls, const float* pos,
                const float* ln_g, const float* ln_b, float* x, void* workspace, int NB, int R, int p, int D) {
  PEVIT_REQUIRE(R % p == 0 && D % 128 == 0 && D <= 128 * FIN_MAXV, "patch_embed: unsupported R=%d p=%d D=%d", R, p, D);
  const int G = R / p, L = G * G + 1, K = 3 * p * p, Kpad = (K + 7) / 8 * 8;
  const int rows = NB * G * G;
  bf16* patches = static_cast<bf16*>(workspace);
  float* emb = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) +
                                        ((static_cast<size_t>(rows) * Kpad * sizeof(bf16) + 255) & ~size_t(255)));
  {
    ProfScope prof(s, PC_STEM);
    // 16-byte loads need p % 4 == 0, R % 4 == 0 and an aligned image base
    const bool vec4 = p % 4 == 0 && R % 4 == 0 && (reinterpret_cast<uintptr_t>(img) & 15) == 0;
    if (vec4)
      PEVIT_CHECK_CUDA(launch_kernel(im2col_kernel<true>, dim3(rows), dim3(256), 0, s, 1, img, patches, NB, R, p, G, K, Kpad));
    else
      PEVIT_CHECK_CUDA(launch_kernel(im2col_kernel<false>, dim3(rows), dim3(256), 0, s, 1, img, patches, NB, R, p, G, K, Kpad));
    PEVIT_CHECK_LAUNCH();
  }
  GemmEpilogue ep;
  ep.out_f32 = emb;
  ep.ld_out = D;
  prof_set_tag(PC_GEMM_STEM);
  int rc = gemm_tn(s, patches, Kpad, w_patch, Kpad, rows, D, Kpad, EPI_F32, ep);
  if (rc != 0) return rc;
  {
    ProfScope prof(s, PC_STEM);
    const int grid = (L * NB + FIN_THREADS / 32 - 1) / (FIN_THREADS / 32);
    PEVIT_CHECK_CUDA(launch_kernel(embed_finalize_kernel, dim3(grid), dim3(FIN_THREADS), 0, s, 1, emb, cls, pos, ln_g, ln_b, x, NB, L, D));
    PEVIT_CHECK_LAUNCH();
  }
  return 0;
}

}  // namespace pevit
