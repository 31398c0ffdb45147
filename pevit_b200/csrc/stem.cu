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
// Pixel formats the stem reads (pevit_pixel_dtype): fp32 (the reference's data loader output), bf16 (the same values
// already rounded: identical patches, half the H2D bytes), raw uint8 with torchvision's ToTensor + Normalize applied
// here in fp32 with the host's operation order ((u8 / 255 - mean) / std, IEEE division: bit-identical to the host).
struct PixelNorm { float mean[3], std[3]; };
template <typename T> struct Px;
template <> struct Px<float> {
  using Vec4 = float4;
  static __device__ __forceinline__ float get(float v, int, const PixelNorm&) { return v; }
  static __device__ __forceinline__ void get4(const Vec4& v, int, const PixelNorm&, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
};
template <> struct Px<bf16> {
  using Vec4 = uint2;
  static __device__ __forceinline__ float get(bf16 v, int, const PixelNorm&) { return __bfloat162float(v); }
  static __device__ __forceinline__ void get4(const Vec4& v, int, const PixelNorm&, float (&o)[4]) {
    o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
    o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
  }
};
template <> struct Px<uint8_t> {
  using Vec4 = uchar4;
  static __device__ __forceinline__ float get(uint8_t v, int c, const PixelNorm& n) {
    return __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.f), n.mean[c]), n.std[c]);
  }
  static __device__ __forceinline__ void get4(const Vec4& v, int c, const PixelNorm& n, float (&o)[4]) {
    o[0] = get(v.x, c, n); o[1] = get(v.y, c, n); o[2] = get(v.z, c, n); o[3] = get(v.w, c, n);
  }
};

template <bool kVec4, typename T>
__global__ void im2col_kernel(const T* __restrict__ img, bf16* __restrict__ patches, int NB, int R, int p, int G,
                              int K, int Kpad, PixelNorm norm) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;  // (n, gy, gx)
  const int n = row / (G * G), g = row - n * G * G;
  const int gy = g / G, gx = g - gy * G;
  const T* src = img + (static_cast<size_t>(n) * 3 * R + gy * p) * R + gx * p;
  bf16* dst = patches + static_cast<size_t>(row) * Kpad;
  const int pp = p * p;
  if constexpr (kVec4) {
    const int p4 = p >> 2;
    for (int q = threadIdx.x; q < (K >> 2); q += blockDim.x) {  // q = (c, i, j/4)
      const int ci = q / p4, j4 = q - ci * p4;
      const int c = ci / p, i = ci - c * p;
      using Vec4 = typename Px<T>::Vec4;
      const Vec4 raw = __ldg(reinterpret_cast<const Vec4*>(src + (static_cast<size_t>(c) * R + i) * R) + j4);
      float v[4];
      Px<T>::get4(raw, c, norm, v);
      *reinterpret_cast<uint2*>(dst + 4 * q) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
    }
    for (int k = K + threadIdx.x; k < Kpad; k += blockDim.x) dst[k] = __float2bfloat16(0.f);
  } else {
    for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
      float v = 0.f;
      if (k < K) {
        const int c = k / pp, rem = k - c * pp;
        const int i = rem / p, j = rem - i * p;
        v = Px<T>::get(__ldg(src + (static_cast<size_t>(c) * R + i) * R + j), c, norm);
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

template <typename T>
static int launch_im2col(cudaStream_t s, const T* img, bf16* patches, int NB, int R, int p, int G, int K, int Kpad,
                         int rows, const PixelNorm& norm) {
  // vector loads of 4 pixels need p % 4 == 0, R % 4 == 0 and a base aligned to 4 pixels
  const bool vec4 = p % 4 == 0 && R % 4 == 0 && (reinterpret_cast<uintptr_t>(img) & (4 * sizeof(T) - 1)) == 0;
  if (vec4)
    PEVIT_CHECK_CUDA(launch_kernel(im2col_kernel<true, T>, dim3(rows), dim3(256), 0, s, 1, img, patches, NB, R, p, G, K, Kpad, norm));
  else
    PEVIT_CHECK_CUDA(launch_kernel(im2col_kernel<false, T>, dim3(rows), dim3(256), 0, s, 1, img, patches, NB, R, p, G, K, Kpad, norm));
  PEVIT_CHECK_LAUNCH();
  return 0;
}

int patch_embed(cudaStream_t s, const void* img, int px_dtype, const float* mean, const float* stdv, const bf16* w_patch,
                const float* cls, const float* pos, const float* ln_g, const float* ln_b, float* x, void* workspace,
                int NB, int R, int p, int D) {
  PEVIT_REQUIRE(px_dtype >= 0 && px_dtype <= 2, "patch_embed: unknown pixel dtype %d", px_dtype);
  PEVIT_REQUIRE(px_dtype != 2 || (mean != nullptr && stdv != nullptr), "patch_embed: uint8 pixels need mean / std");
  PEVIT_REQUIRE(R % p == 0 && D % 128 == 0 && D <= 128 * FIN_MAXV, "patch_embed: unsupported R=%d p=%d D=%d", R, p, D);
  const int G = R / p, L = G * G + 1, K = 3 * p * p, Kpad = (K + 7) / 8 * 8;
  const int rows = NB * G * G;
  bf16* patches = static_cast<bf16*>(workspace);
  float* emb = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) +
                                        ((static_cast<size_t>(rows) * Kpad * sizeof(bf16) + 255) & ~size_t(255)));
  {
    ProfScope prof(s, PC_STEM);
    PixelNorm norm{{0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}};
    if (px_dtype == 2)
      for (int c = 0; c < 3; ++c) { norm.mean[c] = mean[c]; norm.std[c] = stdv[c]; }
    int rc = 0;
    if (px_dtype == 0) rc = launch_im2col(s, static_cast<const float*>(img), patches, NB, R, p, G, K, Kpad, rows, norm);
    else if (px_dtype == 1) rc = launch_im2col(s, static_cast<const bf16*>(img), patches, NB, R, p, G, K, Kpad, rows, norm);
    else rc = launch_im2col(s, static_cast<const uint8_t*>(img), patches, NB, R, p, G, K, Kpad, rows, norm);
    if (rc != 0) return rc;
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
