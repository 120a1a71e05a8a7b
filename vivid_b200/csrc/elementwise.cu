// Fused elementwise passes over the NHWC residual stream (fp32) — HBM-bound, 16-byte
// vector accesses, one warp per pixel so the per-pixel channel reductions are warp shuffles.
// Reference ops: normalize(dim=1) pixel-norm (training/models.py:171, 37-42), resample up/down
// (:48-61), mp_silu (:66-67), mp_cat (:78-84), MPFourier + embedding linears (:96-101, 388-391,
// 175), EDM preconditioning (NVPrecond.forward), Heun/guidance update (generate_images.py:62,
// 93-114) and the uint8 pixel codec (training/encoders.py:58-62).
#include "common.h"
#include "ptx.cuh"

namespace vb {
namespace {

constexpr int kEwThreads = 256;
constexpr int kMaxVec = 8;   // float4 per lane -> up to 1024 channels per pixel

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint2 pack4(float a, float b, float c, float d) {
  return make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
}

// ---- PIXNORM / DOWN_PIXNORM: one warp per output pixel --------------------------------
template <bool DOWN>
__global__ void __launch_bounds__(kEwThreads) pixnorm_kernel(const float* __restrict__ a, float* __restrict__ out_f32,
                                                             __nv_bfloat16* __restrict__ out_bf16,
                                                             __nv_bfloat16* __restrict__ out_silu, long long pixels,
                                                             int H, int W, int C) {
  const long long pix = (static_cast<long long>(blockIdx.x) * kEwThreads + threadIdx.x) >> 5;
  if (pix >= pixels) return;
  const int lane = threadIdx.x & 31;
  const int nvec = C >> 2;
  float4 v[kMaxVec];
  float ss = 0.f;
  if (!DOWN) {
    const float4* src = reinterpret_cast<const float4*>(a + pix * C);
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int i = lane + j * 32;
      if (i < nvec) {
        v[j] = __ldg(src + i);
        ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
      }
    }
  } else {
    // output pixel (n, y, x) <- mean of the 2x2 input patch at (2y, 2x) of a [2H][2W] image
    const int x = static_cast<int>(pix % W);
    const long long t = pix / W;
    const int y = static_cast<int>(t % H);
    const long long n = t / H;
    const long long p00 = (n * 2 * H + 2 * y) * (2 * W) + 2 * x;
    const float4* s0 = reinterpret_cast<const float4*>(a + p00 * C);
    const float4* s1 = reinterpret_cast<const float4*>(a + (p00 + 1) * C);
    const float4* s2 = reinterpret_cast<const float4*>(a + (p00 + 2 * W) * C);
    const float4* s3 = reinterpret_cast<const float4*>(a + (p00 + 2 * W + 1) * C);
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int i = lane + j * 32;
      if (i < nvec) {
        const float4 p = __ldg(s0 + i), q = __ldg(s1 + i), r = __ldg(s2 + i), s = __ldg(s3 + i);
        v[j] = make_float4(0.25f * (p.x + q.x + r.x + s.x), 0.25f * (p.y + q.y + r.y + s.y),
                           0.25f * (p.z + q.z + r.z + s.z), 0.25f * (p.w + q.w + r.w + s.w));
        ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
      }
    }
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / (1e-4f + sqrtf(ss) * rsqrtf(static_cast<float>(C)));
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = lane + j * 32;
    if (i < nvec) {
      const float4 o = make_float4(v[j].x * inv, v[j].y * inv, v[j].z * inv, v[j].w * inv);
      if (out_f32) reinterpret_cast<float4*>(out_f32 + pix * C)[i] = o;
      if (out_bf16) reinterpret_cast<uint2*>(out_bf16 + pix * C)[i] = pack4(o.x, o.y, o.z, o.w);
      if (out_silu)
        reinterpret_cast<uint2*>(out_silu + pix * C)[i] = pack4(mp_silu_f(o.x), mp_silu_f(o.y), mp_silu_f(o.z), mp_silu_f(o.w));
    }
  }
}

// ---- UP: one warp per INPUT pixel, writes the 2x2 output patch ------------------------
__global__ void __launch_bounds__(kEwThreads) up_kernel(const float* __restrict__ a, float* __restrict__ out_f32,
                                                        __nv_bfloat16* __restrict__ out_bf16,
                                                        __nv_bfloat16* __restrict__ out_silu, long long in_pixels, int Hi,
                                                        int Wi, int C) {
  const long long pix = (static_cast<long long>(blockIdx.x) * kEwThreads + threadIdx.x) >> 5;
  if (pix >= in_pixels) return;
  const int lane = threadIdx.x & 31;
  const int nvec = C >> 2;
  const int x = static_cast<int>(pix % Wi);
  const long long t = pix / Wi;
  const int y = static_cast<int>(t % Hi);
  const long long n = t / Hi;
  const long long o00 = (n * 2 * Hi + 2 * y) * (2 * Wi) + 2 * x;
  const long long offs[4] = {o00, o00 + 1, o00 + 2 * Wi, o00 + 2 * Wi + 1};
  const float4* src = reinterpret_cast<const float4*>(a + pix * C);
  for (int i = lane; i < nvec; i += 32) {
    const float4 o = __ldg(src + i);
    const uint2 b = pack4(o.x, o.y, o.z, o.w);
    const uint2 s = pack4(mp_silu_f(o.x), mp_silu_f(o.y), mp_silu_f(o.z), mp_silu_f(o.w));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (out_f32) reinterpret_cast<float4*>(out_f32 + offs[k] * C)[i] = o;
      if (out_bf16) reinterpret_cast<uint2*>(out_bf16 + offs[k] * C)[i] = b;
      if (out_silu) reinterpret_cast<uint2*>(out_silu + offs[k] * C)[i] = s;
    }
  }
}

// ---- CAT / SILU: one warp per pixel ---------------------------------------------------
__global__ void __launch_bounds__(kEwThreads) cat_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                         float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16,
                                                         __nv_bfloat16* __restrict__ out_silu, long long pixels, int ca,
                                                         int cb, float wa, float wb) {
  const long long pix = (static_cast<long long>(blockIdx.x) * kEwThreads + threadIdx.x) >> 5;
  if (pix >= pixels) return;
  const int lane = threadIdx.x & 31;
  const int na = ca >> 2, nb = cb >> 2;
  const int C = ca + cb;
  const float4* sa = reinterpret_cast<const float4*>(a + pix * ca);
  const float4* sb = b ? reinterpret_cast<const float4*>(b + pix * cb) : nullptr;
  for (int i = lane; i < na + nb; i += 32) {
    float4 o;
    if (i < na) {
      o = __ldg(sa + i);
      o = make_float4(o.x * wa, o.y * wa, o.z * wa, o.w * wa);
    } else {
      o = __ldg(sb + (i - na));
      o = make_float4(o.x * wb, o.y * wb, o.z * wb, o.w * wb);
    }
    if (out_f32) reinterpret_cast<float4*>(out_f32 + pix * C)[i] = o;
    if (out_bf16) reinterpret_cast<uint2*>(out_bf16 + pix * C)[i] = pack4(o.x, o.y, o.z, o.w);
    if (out_silu)
      reinterpret_cast<uint2*>(out_silu + pix * C)[i] = pack4(mp_silu_f(o.x), mp_silu_f(o.y), mp_silu_f(o.z), mp_silu_f(o.w));
  }
}

// ---- embedding ------------------------------------------------------------------------
// One block per batch row: Fourier features -> emb_noise (+ emb_label, mp_sum) -> mp_silu.
__global__ void __launch_bounds__(256) emb_kernel(const vb_emb_desc d) {
  extern __shared__ float s_in[];   // [cnoise] fourier, then [label_dim] geometry
  const int b = blockIdx.x;
  const float sigma = d.sigma[d.sigma_n == 1 ? 0 : static_cast<size_t>(b) * d.sigma_stride];
  const float c_noise = logf(sigma) * 0.25f * d.noise_scale;
  for (int c = threadIdx.x; c < d.cnoise; c += blockDim.x)
    s_in[c] = cosf(c_noise * d.freqs[c] + d.phases[c]) * 1.4142135623730951f;
  float* s_geo = s_in + d.cnoise;
  if (d.w_label != nullptr) {
    for (int c = threadIdx.x; c < d.label_dim; c += blockDim.x) {
      float g = 0.f;
      if (d.geom != nullptr) g = d.geom[(d.geom_rows == 1 ? 0 : static_cast<size_t>(b) * d.label_dim) + c] * d.geom_scale;
      s_geo[c] = g;
    }
  }
  __syncthreads();
  const float t = d.label_balance;
  const float inv = rsqrtf((1.f - t) * (1.f - t) + t * t);
  for (int j = threadIdx.x; j < d.cemb; j += blockDim.x) {
    const float* wn = d.w_noise + static_cast<size_t>(j) * d.cnoise;
    float e = 0.f;
    for (int c = 0; c < d.cnoise; ++c) e += wn[c] * s_in[c];
    if (d.w_label != nullptr) {
      const float* wl = d.w_label + static_cast<size_t>(j) * d.label_dim;
      float g = 0.f;
      for (int c = 0; c < d.label_dim; ++c) g += wl[c] * s_geo[c];
      e = (e * (1.f - t) + g * t) * inv;
    }
    // full-precision silu here: this vector modulates every block
    d.emb[static_cast<size_t>(b) * d.cemb + j] = e / (1.f + expf(-e)) * (1.0f / 0.596f);
  }
}
// One warp per modulation channel m: mod[b][m] = Wmod[m] . emb[b] + 1 for every b.
__global__ void __launch_bounds__(256) mod_kernel(const vb_emb_desc d) {
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (m >= d.mod_total) return;
  const int lane = threadIdx.x & 31;
  const float* w = d.w_mod + static_cast<size_t>(m) * d.cemb;
  for (int b = 0; b < d.B; ++b) {
    const float* e = d.emb + static_cast<size_t>(b) * d.cemb;
    float acc = 0.f;
    for (int j = lane; j < d.cemb; j += 32) acc += __ldg(w + j) * e[j];
    acc = warp_sum(acc);
    if (lane == 0) d.mod[static_cast<size_t>(b) * d.mod_total + m] = acc + 1.0f;
  }
}

// ---- preconditioning ------------------------------------------------------------------
__global__ void __launch_bounds__(256) precond_in_kernel(const vb_precond_in_desc d) {
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long hw = static_cast<long long>(d.R) * d.R;
  if (pix >= hw * d.B) return;
  const long long n = pix / hw, s = pix - n * hw;
  float c_in = 1.f;
  if (d.sigma != nullptr) {
    const float sg = d.sigma[d.sigma_n == 1 ? 0 : n * d.sigma_stride];
    c_in = rsqrtf(d.sigma_data * d.sigma_data + sg * sg);
  }
  float ch[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float* xp = d.x + n * d.img_stride + s;
  ch[0] = xp[0] * c_in;
  ch[1] = xp[hw] * c_in;
  ch[2] = xp[2 * hw] * c_in;
  int k = 3;
  if (d.cond != nullptr) {
    const float* cp = d.cond + n * 3 * hw + s;
    const float* np = d.noise ? d.noise + n * 3 * hw + s : nullptr;
    for (int c = 0; c < 3; ++c) ch[3 + c] = cp[c * hw] + (np ? d.noisy_sr * np[c * hw] : 0.f);
    k = 6;
  }
  ch[k] = 1.0f;   // bias-as-channel (training/models.py:394)
  uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(d.out) + pix * d.cpad);
  o[0] = make_uint4(pack_bf16x2(ch[0], ch[1]), pack_bf16x2(ch[2], ch[3]), pack_bf16x2(ch[4], ch[5]), pack_bf16x2(ch[6], ch[7]));
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (int i = 1; i < d.cpad / 8; ++i) o[i] = z;
}

__global__ void __launch_bounds__(256) precond_out_kernel(const vb_precond_out_desc d) {
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long hw = static_cast<long long>(d.R) * d.R;
  if (pix >= hw * d.B) return;
  const long long n = pix / hw, s = pix - n * hw;
  const float sg = d.sigma[d.sigma_n == 1 ? 0 : n * d.sigma_stride];
  const float sd = d.sigma_data;
  const float den = sg * sg + sd * sd;
  const float c_skip = sd * sd / den;
  const float c_out = sg * sd * rsqrtf(den);
  const float4 f = *reinterpret_cast<const float4*>(d.f + pix * d.ldf);
  const float* xp = d.x + n * d.img_stride + s;
  float* op = d.d_out + n * 3 * hw + s;
  op[0] = c_skip * xp[0] + c_out * f.x;
  op[hw] = c_skip * xp[hw] + c_out * f.y;
  op[2 * hw] = c_skip * xp[2 * hw] + c_out * f.z;
}

// ---- Heun + guidance ------------------------------------------------------------------
__global__ void __launch_bounds__(256) heun_kernel(const vb_heun_desc d) {
  const long long i4 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n4 = d.n >> 2;
  const float w = d.guidance, th = d.t_hat, tn = d.t_next;
  auto step = [&](float dn, float dg, float xh, float& dc, float& xn) {
    const float D = d.d_gnet ? dg + w * (dn - dg) : dn;     // lerp(D_g, D_n, w)
    if (d.phase == 0) {
      dc = (xh - D) / th;
      xn = xh + (tn - th) * dc;
    } else {
      const float dp = (xn - D) / tn;
      xn = xh + (tn - th) * (0.5f * dc + 0.5f * dp);
    }
  };
  if (i4 < n4) {
    const float4 dn = reinterpret_cast<const float4*>(d.d_net)[i4];
    const float4 dg = d.d_gnet ? reinterpret_cast<const float4*>(d.d_gnet)[i4] : dn;
    const float4 xh = reinterpret_cast<const float4*>(d.x_hat)[i4];
    float4 dc = d.phase == 0 ? make_float4(0, 0, 0, 0) : reinterpret_cast<const float4*>(d.d_cur)[i4];
    float4 xn = d.phase == 0 ? make_float4(0, 0, 0, 0) : reinterpret_cast<const float4*>(d.x_next)[i4];
    step(dn.x, dg.x, xh.x, dc.x, xn.x);
    step(dn.y, dg.y, xh.y, dc.y, xn.y);
    step(dn.z, dg.z, xh.z, dc.z, xn.z);
    step(dn.w, dg.w, xh.w, dc.w, xn.w);
    if (d.phase == 0) reinterpret_cast<float4*>(d.d_cur)[i4] = dc;
    reinterpret_cast<float4*>(d.x_next)[i4] = xn;
  }
  if (i4 == 0) {   // tail (n not a multiple of 4)
    for (long long i = n4 << 2; i < d.n; ++i) {
      float dc = d.phase == 0 ? 0.f : d.d_cur[i], xn = d.phase == 0 ? 0.f : d.x_next[i];
      step(d.d_net[i], d.d_gnet ? d.d_gnet[i] : 0.f, d.x_hat[i], dc, xn);
      if (d.phase == 0) d.d_cur[i] = dc;
      d.x_next[i] = xn;
    }
  }
}

__global__ void encode_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = static_cast<float>(src[i]) / 127.5f - 1.0f;
}
__global__ void decode_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    // (x*127.5+128).clip(0,255).to(uint8): the cast truncates toward zero
    const float v = fminf(fmaxf(src[i] * 127.5f + 128.0f, 0.f), 255.f);
    dst[i] = static_cast<uint8_t>(v);
  }
}

inline unsigned warp_grid(long long warps) { return static_cast<unsigned>((warps * 32 + kEwThreads - 1) / kEwThreads); }

}  // namespace

int eltwise_launch(const vb_ew_desc* d, cudaStream_t s) {
  VB_REQUIRE(d != nullptr && d->a != nullptr, "vb_eltwise: null input");
  VB_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0, "vb_eltwise: empty extent");
  VB_REQUIRE(d->ca > 0 && d->ca % 4 == 0 && d->ca <= kMaxVec * 128, "vb_eltwise: channels %d must be a multiple of 4, <= %d",
             d->ca, kMaxVec * 128);
  VB_REQUIRE(d->out_f32 || d->out_bf16 || d->out_silu, "vb_eltwise: no output");
  const long long pixels = static_cast<long long>(d->B) * d->H * d->W;
  __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(d->out_bf16);
  __nv_bfloat16* os = static_cast<__nv_bfloat16*>(d->out_silu);
  switch (d->kind) {
    case VB_EW_PIXNORM:
      pixnorm_kernel<false><<<warp_grid(pixels), kEwThreads, 0, s>>>(d->a, d->out_f32, ob, os, pixels, d->H, d->W, d->ca);
      break;
    case VB_EW_DOWN_PIXNORM:
      pixnorm_kernel<true><<<warp_grid(pixels), kEwThreads, 0, s>>>(d->a, d->out_f32, ob, os, pixels, d->H, d->W, d->ca);
      break;
    case VB_EW_UP: {
      VB_REQUIRE(d->H % 2 == 0 && d->W % 2 == 0, "vb_eltwise: UP needs even output extent");
      const long long in_pixels = pixels / 4;
      up_kernel<<<warp_grid(in_pixels), kEwThreads, 0, s>>>(d->a, d->out_f32, ob, os, in_pixels, d->H / 2, d->W / 2, d->ca);
      break;
    }
    case VB_EW_CAT:
      VB_REQUIRE(d->b != nullptr && d->cb > 0 && d->cb % 4 == 0, "vb_eltwise: CAT needs b with cb %% 4 == 0");
      cat_kernel<<<warp_grid(pixels), kEwThreads, 0, s>>>(d->a, d->b, d->out_f32, ob, os, pixels, d->ca, d->cb, d->wa, d->wb);
      break;
    case VB_EW_SILU:
      cat_kernel<<<warp_grid(pixels), kEwThreads, 0, s>>>(d->a, nullptr, d->out_f32, ob, os, pixels, d->ca, 0, 1.0f, 1.0f);
      break;
    default:
      VB_REQUIRE(false, "vb_eltwise: unknown kind %d", d->kind);
  }
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

int embed_launch(const vb_emb_desc* d, cudaStream_t s) {
  VB_REQUIRE(d != nullptr && d->sigma && d->freqs && d->phases && d->w_noise && d->emb, "vb_embed: null argument");
  VB_REQUIRE(d->B > 0 && d->cnoise > 0 && d->cemb > 0, "vb_embed: empty problem");
  VB_REQUIRE(d->mod_total == 0 || (d->w_mod && d->mod), "vb_embed: w_mod/mod missing");
  const size_t smem = sizeof(float) * (d->cnoise + (d->w_label ? d->label_dim : 0));
  emb_kernel<<<d->B, 256, smem, s>>>(*d);
  VB_CHECK_CUDA(cudaGetLastError());
  if (d->mod_total > 0) {
    mod_kernel<<<(d->mod_total * 32 + 255) / 256, 256, 0, s>>>(*d);
    VB_CHECK_CUDA(cudaGetLastError());
  }
  return VB_OK;
}

int precond_in_launch(const vb_precond_in_desc* d, cudaStream_t s) {
  VB_REQUIRE(d != nullptr && d->x && d->out, "vb_precond_in: null tensor");
  VB_REQUIRE(d->B > 0 && d->R > 0 && d->cpad >= 8 && d->cpad % 8 == 0, "vb_precond_in: bad extent");
  const long long pixels = static_cast<long long>(d->B) * d->R * d->R;
  precond_in_kernel<<<static_cast<unsigned>((pixels + 255) / 256), 256, 0, s>>>(*d);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

int precond_out_launch(const vb_precond_out_desc* d, cudaStream_t s) {
  VB_REQUIRE(d != nullptr && d->x && d->f && d->sigma && d->d_out, "vb_precond_out: null tensor");
  VB_REQUIRE(d->B > 0 && d->R > 0 && d->ldf >= 4 && d->ldf % 4 == 0, "vb_precond_out: bad extent");
  const long long pixels = static_cast<long long>(d->B) * d->R * d->R;
  precond_out_kernel<<<static_cast<unsigned>((pixels + 255) / 256), 256, 0, s>>>(*d);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

int heun_launch(const vb_heun_desc* d, cudaStream_t s) {
  VB_REQUIRE(d != nullptr && d->d_net && d->x_hat && d->d_cur && d->x_next, "vb_heun: null tensor");
  VB_REQUIRE(d->n > 0 && (d->phase == 0 || d->phase == 1), "vb_heun: bad n/phase");
  const long long n4 = (d->n >> 2) > 0 ? (d->n >> 2) : 1;
  heun_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, s>>>(*d);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

}  // namespace vb

extern "C" int vb_eltwise(const vb_ew_desc* d, void* stream) { return vb::eltwise_launch(d, static_cast<cudaStream_t>(stream)); }
extern "C" int vb_embed(const vb_emb_desc* d, void* stream) { return vb::embed_launch(d, static_cast<cudaStream_t>(stream)); }
extern "C" int vb_precond_in(const vb_precond_in_desc* d, void* stream) {
  return vb::precond_in_launch(d, static_cast<cudaStream_t>(stream));
}
extern "C" int vb_precond_out(const vb_precond_out_desc* d, void* stream) {
  return vb::precond_out_launch(d, static_cast<cudaStream_t>(stream));
}
extern "C" int vb_heun(const vb_heun_desc* d, void* stream) { return vb::heun_launch(d, static_cast<cudaStream_t>(stream)); }
extern "C" int vb_encode_u8(const uint8_t* src, float* dst, int64_t n, void* stream) {
  VB_REQUIRE(src && dst && n > 0, "vb_encode_u8: bad argument");
  vb::encode_u8_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}
extern "C" int vb_decode_u8(const float* src, uint8_t* dst, int64_t n, void* stream) {
  VB_REQUIRE(src && dst && n > 0, "vb_decode_u8: bad argument");
  vb::decode_u8_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}
