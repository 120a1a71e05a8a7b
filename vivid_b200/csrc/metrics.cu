// Statistics leg of `calculate_metrics.py gen` and the inter-stage image resize (N3).
//
//  vb_stats_update   cum_mu += sum_n f[n,:],  cum_sigma += f^T f  in fp64 (calculate_metrics.py:158-172: the reference does
//                    `features.to(float64)`, `.sum(0)` and `features.T @ features` per batch, four times with the joint
//                    variants); here one pass over the upper-triangular 64x64 blocks, mirrored on write, with the joint
//                    [features | src_features] concatenation read in place.
//  vb_psnr_u8        per-image PSNR of uint8 images against a target (calculate_metrics.py:148), fp64 statistics.
//  vb_resize         torch.nn.functional.interpolate(mode="bilinear", align_corners=False [, antialias=True]) for fp32 NCHW
//                    images: the x4 inter-stage upscale (generate_images.py:322) and the x1/4 anti-aliased low-res
//                    conditioning of the SR-only path (generate_images.py:282-283).
// HBM-bound byte/float work: coalesced rows, fp64 accumulation in registers, no atomics (deterministic sums).
#include "common.h"
#include "ptx.cuh"

namespace vb {
namespace {

constexpr int kStTile = 64;     // output block edge
constexpr int kStK = 16;        // samples per shared-memory step

template <typename T>
__device__ __forceinline__ double to_f64(T v);
template <>
__device__ __forceinline__ double to_f64<float>(float v) { return static_cast<double>(v); }
template <>
__device__ __forceinline__ double to_f64<double>(double v) { return v; }
template <>
__device__ __forceinline__ double to_f64<__half>(__half v) { return static_cast<double>(__half2float(v)); }
template <>
__device__ __forceinline__ double to_f64<__nv_bfloat16>(__nv_bfloat16 v) { return static_cast<double>(__bfloat162float(v)); }

// Feature column c of sample n of the (virtually concatenated) matrix [x1 | x2].
template <typename T>
__device__ __forceinline__ double feat_at(const vb_stats_desc& d, int n, int c) {
  if (c < d.f1) return to_f64(static_cast<const T*>(d.feat)[static_cast<long long>(n) * d.ld1 + c]);
  return to_f64(static_cast<const T*>(d.feat2)[static_cast<long long>(n) * d.ld2 + (c - d.f1)]);
}

// One CTA (16 x 16 threads, 4 x 4 outputs each) per upper-triangular 64 x 64 block of cum_sigma.
template <typename T>
__global__ void __launch_bounds__(256) stats_sigma_kernel(const vb_stats_desc d) {
  __shared__ double sa[kStK][kStTile];
  __shared__ double sb[kStK][kStTile];
  pdl_grid_sync();
  const int F = d.f1 + d.f2;
  // decode the linear block index into (bi <= bj)
  const int nb = (F + kStTile - 1) / kStTile;
  int bi = 0, rem = blockIdx.x;
  while (rem >= nb - bi) {
    rem -= nb - bi;
    ++bi;
  }
  const int bj = bi + rem;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int n0 = 0; n0 < d.n; n0 += kStK) {
    for (int e = threadIdx.x; e < kStK * kStTile; e += 256) {
      const int k = e / kStTile, c = e % kStTile;
      const int n = n0 + k;
      const int ci = bi * kStTile + c, cj = bj * kStTile + c;
      sa[k][c] = (n < d.n && ci < F) ? feat_at<T>(d, n, ci) : 0.0;
      sb[k][c] = (n < d.n && cj < F) ? feat_at<T>(d, n, cj) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kStK; ++k) {
      double ra[4], rb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) ra[a] = sa[k][ty * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) rb[b] = sb[k][tx * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(ra[a], rb[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = bi * kStTile + ty * 4 + a;
    if (i >= F) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = bj * kStTile + tx * 4 + b;
      if (j >= F) continue;
      d.cum_sigma[static_cast<long long>(i) * F + j] += acc[a][b];
      if (bi != bj) d.cum_sigma[static_cast<long long>(j) * F + i] += acc[a][b];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) stats_mu_kernel(const vb_stats_desc d) {
  pdl_grid_sync();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= d.f1 + d.f2) return;
  double s = 0.0;
  for (int n = 0; n < d.n; ++n) s += feat_at<T>(d, n, c);
  d.cum_mu[c] += s;
}

template <typename T>
int stats_launch_t(const vb_stats_desc* d, cudaStream_t s) {
  const int F = d->f1 + d->f2;
  const int nb = (F + kStTile - 1) / kStTile;
  VB_CHECK_CUDA(launch_pdl(stats_mu_kernel<T>, dim3((F + 255) / 256), dim3(256), 0, s, *d));
  VB_CHECK_CUDA(launch_pdl(stats_sigma_kernel<T>, dim3(nb * (nb + 1) / 2), dim3(256), 0, s, *d));
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------ PSNR
template <typename T>
__device__ __forceinline__ float tgt_val(const void* p, long long i);
template <>
__device__ __forceinline__ float tgt_val<uint8_t>(const void* p, long long i) {
  return static_cast<float>(static_cast<const uint8_t*>(p)[i]);
}
template <>
__device__ __forceinline__ float tgt_val<float>(const void* p, long long i) { return static_cast<const float*>(p)[i]; }

// One CTA per image: differences and their squares in fp64 (exact for uint8 / fp32 inputs),
// fixed-shape tree reduction, psnr = 10 log10(255^2 / mse).
template <typename T>
__global__ void __launch_bounds__(256) psnr_kernel(const uint8_t* __restrict__ img, const void* __restrict__ tgt,
                                                   long long per_image, long long tgt_stride, double* __restrict__ out) {
  __shared__ double red[256];
  pdl_grid_sync();
  const uint8_t* x = img + static_cast<long long>(blockIdx.x) * per_image;
  const long long t0 = static_cast<long long>(blockIdx.x) * tgt_stride;
  double s = 0.0;
  for (long long i = threadIdx.x; i < per_image; i += 256) {
    const double df = static_cast<double>(x[i]) - static_cast<double>(tgt_val<T>(tgt, t0 + i));
    s = fma(df, df, s);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = 10.0 * log10(255.0 * 255.0 / (red[0] / static_cast<double>(per_image)));
}

__global__ void psnr_sum_kernel(const double* __restrict__ v, int n, double* __restrict__ cum) {
  pdl_grid_sync();
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += v[i];      // image order: the same sum on every run
  cum[0] += s;
}

// ------------------------------------------------------------------------------------------------ resize
// Separable triangle-filter resize, the arithmetic of ATen's upsample_bilinear2d (align_corners=False) and of its
// anti-aliased variant (_upsample_bilinear2d_aa): source coordinate (o + 0.5) * scale - 0.5; plain mode interpolates the two
// neighbours (clamped), anti-alias mode (down-scaling) weights every tap within `scale` of the centre and normalises.
struct Taps {
  int first, count;
  float w[20];
};

// Anti-aliased taps of output index o (ATen _compute_indices_min_size_weights_aa with the triangle filter).
__device__ __forceinline__ Taps make_taps_aa(int o, int in_size, float scale) {
  Taps t;
  const float support = scale >= 1.0f ? scale : 1.0f;      // interp_size 2 -> (2 / 2) * scale
  const float inv = scale >= 1.0f ? 1.0f / scale : 1.0f;
  const float center = scale * (static_cast<float>(o) + 0.5f);
  const int lo = max(static_cast<int>(center - support + 0.5f), 0);
  const int hi = min(static_cast<int>(center + support + 0.5f), in_size);
  t.first = lo;
  t.count = min(hi - lo, 20);
  float tot = 0.f;
  for (int j = 0; j < t.count; ++j) {
    const float x = fabsf((static_cast<float>(j + lo) - center + 0.5f) * inv);
    t.w[j] = x < 1.0f ? 1.0f - x : 0.f;
    tot += t.w[j];
  }
  for (int j = 0; j < t.count; ++j) t.w[j] /= tot;
  return t;
}

// Plain bilinear: index of the first neighbour, step to the second (0 at the border) and its weight.
__device__ __forceinline__ void lerp_taps(int o, int in_size, float scale, int& i0, int& step, float& l1) {
  float src = (static_cast<float>(o) + 0.5f) * scale - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = min(static_cast<int>(src), in_size - 1);
  step = i0 < in_size - 1 ? 1 : 0;
  l1 = fminf(fmaxf(src - static_cast<float>(i0), 0.f), 1.0f);
}

// One thread per output pixel; vertical taps outer, horizontal inner (ATen resizes horizontally first, then vertically:
// the sums below use the same grouping, sum_y wy * (sum_x wx * v)).
__global__ void __launch_bounds__(256) resize_kernel(const float* __restrict__ src, float* __restrict__ dst, int planes, int hi,
                                                     int wi, int ho, int wo, float sy, float sx, int antialias) {
  pdl_grid_sync();
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  const long long total = static_cast<long long>(planes) * ho * wo;
  if (idx >= total) return;
  const int ox = static_cast<int>(idx % wo);
  const int oy = static_cast<int>((idx / wo) % ho);
  const long long pl = idx / (static_cast<long long>(wo) * ho);
  const float* base = src + pl * hi * wi;
  float acc = 0.f;
  if (antialias) {
    const Taps ty = make_taps_aa(oy, hi, sy);
    const Taps tx = make_taps_aa(ox, wi, sx);
    for (int a = 0; a < ty.count; ++a) {
      const float* row = base + static_cast<long long>(ty.first + a) * wi + tx.first;
      float h = 0.f;
      for (int b = 0; b < tx.count; ++b) h += tx.w[b] * row[b];
      acc += ty.w[a] * h;
    }
  } else {
    // upsample_bilinear2d: h0 * (w0 v00 + w1 v01) + h1 * (w0 v10 + w1 v11)
    int y0, ys, x0, xs;
    float hy, hx;
    lerp_taps(oy, hi, sy, y0, ys, hy);
    lerp_taps(ox, wi, sx, x0, xs, hx);
    const float* r0 = base + static_cast<long long>(y0) * wi + x0;
    const float* r1 = r0 + static_cast<long long>(ys) * wi;
    acc = (1.0f - hy) * ((1.0f - hx) * r0[0] + hx * r0[xs]) + hy * ((1.0f - hx) * r1[0] + hx * r1[xs]);
  }
  dst[idx] = acc;
}

}  // namespace
}  // namespace vb

extern "C" int vb_stats_update(const vb_stats_desc* d, void* stream) {
  VB_REQUIRE(d != nullptr && d->feat && d->cum_mu && d->cum_sigma, "vb_stats_update: null argument");
  VB_REQUIRE(d->n > 0 && d->f1 > 0 && d->f2 >= 0 && d->ld1 >= d->f1, "vb_stats_update: bad extent");
  VB_REQUIRE(d->f2 == 0 || (d->feat2 != nullptr && d->ld2 >= d->f2), "vb_stats_update: second feature block missing");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (d->dtype) {
    case VB_F32: return vb::stats_launch_t<float>(d, s);
    case VB_F16: return vb::stats_launch_t<__half>(d, s);
    case VB_BF16: return vb::stats_launch_t<__nv_bfloat16>(d, s);
    case VB_F64: return vb::stats_launch_t<double>(d, s);
    default: break;
  }
  VB_REQUIRE(false, "vb_stats_update: unsupported dtype %d", d->dtype);
}

extern "C" int vb_psnr_u8(const uint8_t* images, const void* tgt, int32_t tgt_dtype, int32_t n, int64_t per_image,
                          int64_t tgt_image_stride, double* psnr_out, double* cum_sum, void* stream) {
  VB_REQUIRE(images && tgt && psnr_out, "vb_psnr_u8: null argument");
  VB_REQUIRE(n > 0 && per_image > 0 && tgt_image_stride >= per_image, "vb_psnr_u8: bad extent");
  VB_REQUIRE(tgt_dtype == VB_U8 || tgt_dtype == VB_F32, "vb_psnr_u8: target must be uint8 or fp32");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (tgt_dtype == VB_U8)
    VB_CHECK_CUDA(vb::launch_pdl(vb::psnr_kernel<uint8_t>, dim3(n), dim3(256), 0, s, images, tgt, per_image, tgt_image_stride, psnr_out));
  else
    VB_CHECK_CUDA(vb::launch_pdl(vb::psnr_kernel<float>, dim3(n), dim3(256), 0, s, images, tgt, per_image, tgt_image_stride, psnr_out));
  if (cum_sum != nullptr)
    VB_CHECK_CUDA(vb::launch_pdl(vb::psnr_sum_kernel, dim3(1), dim3(1), 0, s, static_cast<const double*>(psnr_out), n, cum_sum));
  return VB_OK;
}

extern "C" int vb_resize(const float* src, float* dst, int32_t planes, int32_t h_in, int32_t w_in, int32_t h_out, int32_t w_out,
                         int32_t antialias, void* stream) {
  VB_REQUIRE(src && dst, "vb_resize: null tensor");
  VB_REQUIRE(planes > 0 && h_in > 0 && w_in > 0 && h_out > 0 && w_out > 0, "vb_resize: empty extent");
  const float sy = static_cast<float>(h_in) / static_cast<float>(h_out);
  const float sx = static_cast<float>(w_in) / static_cast<float>(w_out);
  VB_REQUIRE(!antialias || (sy <= 8.0f && sx <= 8.0f), "vb_resize: anti-aliased down-scaling is limited to a factor of 8");
  const long long total = static_cast<long long>(planes) * h_out * w_out;
  VB_CHECK_CUDA(vb::launch_pdl(vb::resize_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0,
                               static_cast<cudaStream_t>(stream), src, dst, planes, h_in, w_in, h_out, w_out, sy, sx, antialias));
  return VB_OK;
}
