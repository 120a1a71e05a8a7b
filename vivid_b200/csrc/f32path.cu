// fp32 validation path (NVPrecond(use_fp16=False) / forward(force_fp32=True); reference training/models.py:632,697).
//
// The production path keeps activations in fp16 like the reference's own reduced-precision mode; the north-star parity
// bound for "fp32 mode" (rel-L2 <= 1e-4 against the reference's fp32 path) needs real fp32 arithmetic, which the tensor
// cores do not offer (kind::tf32 keeps 10 mantissa bits).  This file is that mode: plain CUDA-core kernels, NHWC fp32
// activations, unfused — one kernel per reference op (MPConv's conv2d, normalize, mp_silu, mp_sum, mp_cat, resample,
// qkv normalisation, attention).  It exists to VALIDATE (tests, spot checks of a trained net), not to be fast:
// ~5 TFLOP/s.  Weight normalisation, the embedding MLP and the output preconditioning reuse the production kernels,
// which are fp32 already (weights.cu, elementwise.cu).
#include "common.h"
#include "ptx.cuh"

namespace vb {
namespace {

// ------------------------------------------------------------------------------------------------ convolution
// out[p][co] = sum_{tap,ci} x[pixel(p) + tap][ci] * w[co][ci*taps + tap]   (3x3 same-pad or 1x1; w as OIHW)
// 64 pixels x 64 output channels per CTA, 4 x 4 per thread, K in steps of 16 input channels per tap.
constexpr int kCvM = 64, kCvN = 64, kCvK = 16;

__global__ void __launch_bounds__(256) conv_f32_kernel(const vb_f32_conv_desc d) {
  __shared__ float sa[kCvK][kCvM + 4];
  __shared__ float sb[kCvK][kCvN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long P = static_cast<long long>(d.B) * d.H * d.W;
  const long long p0 = static_cast<long long>(blockIdx.x) * kCvM;
  const int n0 = blockIdx.y * kCvN;
  const int K = d.cin * d.taps;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // this thread's load assignments: A: pixel (tid / 4), 4 channels (tid % 4)*4..; B: cout (tid / 4), 4 channels
  const int lp = threadIdx.x >> 2, lc = (threadIdx.x & 3) * 4;
  const long long pl = p0 + lp;
  const bool pv = pl < P;
  const int px = pv ? static_cast<int>(pl % d.W) : 0;
  const int py = pv ? static_cast<int>((pl / d.W) % d.H) : 0;
  const long long pimg = pv ? pl - static_cast<long long>(py) * d.W - px : 0;     // first pixel of the image
  const int co_l = n0 + lp;
  for (int tap = 0; tap < d.taps; ++tap) {
    const int dy = d.taps == 9 ? tap / 3 - 1 : 0, dx = d.taps == 9 ? tap % 3 - 1 : 0;
    const int yy = py + dy, xx = px + dx;
    const bool inside = pv && yy >= 0 && yy < d.H && xx >= 0 && xx < d.W;
    const float* arow = d.x + (pimg + static_cast<long long>(yy) * d.W + xx) * d.cin;
    for (int c0 = 0; c0 < d.cin; c0 += kCvK) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + lc + j;
        sa[lc + j][lp] = (inside && c < d.cin) ? arow[c] : 0.f;
        sb[lc + j][lp] = (co_l < d.cout && c < d.cin) ? d.w[static_cast<long long>(co_l) * K + static_cast<long long>(c) * d.taps + tap] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kCvK; ++k) {
        float ra[4], rb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ra[i] = sa[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) rb[j] = sb[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ra[i], rb[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long p = p0 + ty * 4 + i;
    if (p >= P) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co < d.cout) d.out[p * d.ldo + co] = acc[i][j];
    }
  }
}

// ------------------------------------------------------------------------------------------------ pointwise ops
__device__ __forceinline__ float mp_silu_f(float x) { return x / (1.0f + expf(-x)) * (1.0f / 0.596f); }

// ACT: one warp per pixel.  out = [mp_silu]( [normalize_C](a) [* mod[b][c]] )       (models.py:37-42,66-67,175-176)
__global__ void __launch_bounds__(256) act_f32_kernel(const vb_f32_op_desc d) {
  const long long P = static_cast<long long>(d.B) * d.H * d.W;
  const long long p = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= P) return;
  const float* a = d.a + p * d.ca;
  float scale = 1.0f;
  if (d.flags & VB_F32_NORM) {
    float ss = 0.f;
    for (int c = lane; c < d.ca; c += 32) ss = fmaf(a[c], a[c], ss);
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    scale = 1.0f / (1e-4f + sqrtf(ss) * rsqrtf(static_cast<float>(d.ca)));
  }
  const float* m = (d.flags & VB_F32_MOD) ? d.mod + (p / (static_cast<long long>(d.H) * d.W)) * d.mod_stride : nullptr;
  for (int c = lane; c < d.ca; c += 32) {
    float v = a[c] * scale;
    if (m) v *= m[c];
    if (d.flags & VB_F32_SILU) v = mp_silu_f(v);
    d.out[p * d.ca + c] = v;
  }
}

// SUM: out = clip(wa * a + wb * b)  (mp_sum with wa = (1-t)/n, wb = t/n, models.py:72-73; clip :204-205); b may be NULL.
__global__ void __launch_bounds__(256) sum_f32_kernel(const vb_f32_op_desc d, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  float v = d.wa * d.a[i];
  if (d.b) v = fmaf(d.wb, d.b[i], v);
  if (d.clip > 0.f) v = fminf(fmaxf(v, -d.clip), d.clip);
  d.out[i] = v;
}

// CAT: out[p] = [wa * a[p] | wb * b[p]]   (mp_cat, models.py:78-84)
__global__ void __launch_bounds__(256) cat_f32_kernel(const vb_f32_op_desc d, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const int C = d.ca + d.cb;
  const long long p = i / C;
  const int c = static_cast<int>(i - p * C);
  d.out[i] = c < d.ca ? d.wa * d.a[p * d.ca + c] : d.wb * d.b[p * d.cb + (c - d.ca)];
}

// DOWN: 2x2 mean; UP: nearest x2 (resample with f=[1,1], models.py:48-61).  H, W are the OUTPUT extents.
__global__ void __launch_bounds__(256) resample_f32_kernel(const vb_f32_op_desc d, long long n, int up) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const int c = static_cast<int>(i % d.ca);
  const long long p = i / d.ca;
  const int x = static_cast<int>(p % d.W), y = static_cast<int>((p / d.W) % d.H);
  const long long b = p / (static_cast<long long>(d.W) * d.H);
  if (up) {
    const int Hi = d.H / 2, Wi = d.W / 2;
    d.out[i] = d.a[((b * Hi + y / 2) * Wi + x / 2) * d.ca + c];
  } else {
    const int Hi = d.H * 2, Wi = d.W * 2;
    const float* s = d.a + ((b * Hi + 2 * y) * Wi + 2 * x) * d.ca + c;
    d.out[i] = 0.25f * (s[0] + s[d.ca] + s[static_cast<long long>(Wi) * d.ca] + s[static_cast<long long>(Wi + 1) * d.ca]);
  }
}

// QKV: a = 1x1 conv output [P][heads*parts*D] with channel h*parts*D + j*D + dd (de-interleaved by vb_weight_prep);
// normalise over dd per (pixel, head, part) and scatter part j to out_j[(b*heads + h)*seq_j + off_j + seg*HW + s][dd],
// image n = b*seg_div + seg  (models.py:192-193, 283-297).  One warp per (pixel, head, part).
__global__ void __launch_bounds__(256) qkv_f32_kernel(const vb_f32_op_desc d) {
  const long long P = static_cast<long long>(d.B) * d.H * d.W;
  const long long g = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int groups = d.heads * d.parts;
  if (g >= P * groups) return;
  const long long p = g / groups;
  const int hj = static_cast<int>(g - p * groups);
  const int h = hj / d.parts, j = hj - h * d.parts;
  const float* a = d.a + p * (static_cast<long long>(groups) * d.head_dim) + static_cast<long long>(hj) * d.head_dim;
  float ss = 0.f;
  for (int c = lane; c < d.head_dim; c += 32) ss = fmaf(a[c], a[c], ss);
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.0f / (1e-4f + sqrtf(ss) * rsqrtf(static_cast<float>(d.head_dim)));
  const int HW = d.H * d.W;
  const int n = static_cast<int>(p / HW), s = static_cast<int>(p - static_cast<long long>(n) * HW);
  const int b = n / d.seg_div, seg = n - b * d.seg_div;
  float* base = j == 0 ? d.out : (j == 1 ? d.out2 : d.out3);
  const long long tok = (static_cast<long long>(b) * d.heads + h) * d.part_seq[j] + d.part_off[j] + static_cast<long long>(seg) * HW + s;
  for (int c = lane; c < d.head_dim; c += 32) base[tok * d.head_dim + c] = a[c] * inv;
}

// PRECOND_IN: NCHW fp32 image(s) -> NHWC rows [c_in * x (3) | cond + noisy_sr * noise (3, optional) | 1]
// (NVPrecond.forward, snapshot models.py:588-611; the ones channel is UNet.forward's, :394).
__global__ void __launch_bounds__(256) precond_in_f32_kernel(const vb_f32_op_desc d) {
  const long long hw = static_cast<long long>(d.H) * d.W;
  const long long p = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (p >= hw * d.B) return;
  const long long b = p / hw, s = p - b * hw;
  float cin = 1.0f;
  if (d.mod) {      // mod = sigma here
    const float sg = d.mod[b * d.mod_stride];
    cin = rsqrtf(d.wa * d.wa + sg * sg);      // wa = sigma_data
  }
  float* o = d.out + p * d.ca;
  int k = 0;
  for (int c = 0; c < 3; ++c) o[k++] = cin * d.a[b * d.img_stride + c * hw + s];
  if (d.b) {
    for (int c = 0; c < 3; ++c) {
      float v = d.b[b * 3 * hw + c * hw + s];
      if (d.b2) v = fmaf(d.wb, d.b2[b * 3 * hw + c * hw + s], v);      // wb = noisy_sr
      o[k++] = v;
    }
  }
  o[k] = 1.0f;
}

// ------------------------------------------------------------------------------------------------ attention
// y[b][s][h*D + dd] = sum_k softmax_k(q.k / sqrt(D)) v[k][dd]  (models.py:196-198); q [B*heads][sq][D], k/v [B*heads][sk][D];
// zero_keys extra keys with k = v = 0 (the unconditional model's zero features).  One warp per query, two passes.
template <int D>
__global__ void __launch_bounds__(256) attn_f32_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                       const float* __restrict__ v, float* __restrict__ y, int heads, int sq,
                                                       int sk, int zero_keys, long long total) {
  const long long w = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= total) return;
  const long long bh = w / sq;
  const int s = static_cast<int>(w - bh * sq);
  const float* qr = q + (bh * sq + s) * D;
  const float* kb = k + bh * sk * D;
  const float* vb_ = v + bh * sk * D;
  float qv[D];
  const float sc = rsqrtf(static_cast<float>(D));
#pragma unroll
  for (int c = 0; c < D; ++c) qv[c] = qr[c] * sc;
  float mx = zero_keys > 0 ? 0.f : -INFINITY;
  for (int j = lane; j < sk; j += 32) {
    const float* kr = kb + static_cast<long long>(j) * D;
    float l = 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c) l = fmaf(qv[c], kr[c], l);
    mx = fmaxf(mx, l);
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float acc[D];
#pragma unroll
  for (int c = 0; c < D; ++c) acc[c] = 0.f;
  float den = 0.f;
  for (int j = lane; j < sk; j += 32) {
    const float* kr = kb + static_cast<long long>(j) * D;
    const float* vr = vb_ + static_cast<long long>(j) * D;
    float l = 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c) l = fmaf(qv[c], kr[c], l);
    const float pr = expf(l - mx);
    den += pr;
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = fmaf(pr, vr[c], acc[c]);
  }
  for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
  den += static_cast<float>(zero_keys) * expf(-mx);
  const long long b = bh / heads;
  const int h = static_cast<int>(bh - b * heads);
  float* yr = y + (b * sq + s) * (static_cast<long long>(heads) * D) + h * D;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    float t = acc[c];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == (c & 31)) yr[c] = t / den;
  }
}

}  // namespace
}  // namespace vb

extern "C" int vb_f32_conv(const vb_f32_conv_desc* d, void* stream) {
  VB_REQUIRE(d != nullptr && d->x && d->w && d->out, "vb_f32_conv: null tensor");
  VB_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->cin > 0 && d->cout > 0, "vb_f32_conv: empty extent");
  VB_REQUIRE(d->taps == 1 || d->taps == 9, "vb_f32_conv: taps must be 1 or 9");
  VB_REQUIRE(d->ldo >= d->cout, "vb_f32_conv: ldo < cout");
  const long long P = static_cast<long long>(d->B) * d->H * d->W;
  const dim3 grid(static_cast<unsigned>((P + vb::kCvM - 1) / vb::kCvM), (d->cout + vb::kCvN - 1) / vb::kCvN);
  vb::conv_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*d);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_f32_op(const vb_f32_op_desc* d, void* stream) {
  VB_REQUIRE(d != nullptr && d->a && d->out, "vb_f32_op: null tensor");
  VB_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->ca > 0, "vb_f32_op: empty extent");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long P = static_cast<long long>(d->B) * d->H * d->W;
  auto blocks = [](long long n, int per) { return static_cast<unsigned>((n + per - 1) / per); };
  switch (d->kind) {
    case VB_F32_ACT:
      VB_REQUIRE(!(d->flags & VB_F32_MOD) || d->mod, "vb_f32_op: ACT with MOD needs mod");
      vb::act_f32_kernel<<<blocks(P, 8), 256, 0, s>>>(*d);
      break;
    case VB_F32_SUM:
      vb::sum_f32_kernel<<<blocks(P * d->ca, 256), 256, 0, s>>>(*d, P * d->ca);
      break;
    case VB_F32_CAT:
      VB_REQUIRE(d->b && d->cb > 0, "vb_f32_op: CAT needs b");
      vb::cat_f32_kernel<<<blocks(P * (d->ca + d->cb), 256), 256, 0, s>>>(*d, P * (d->ca + d->cb));
      break;
    case VB_F32_DOWN:
    case VB_F32_UP:
      VB_REQUIRE(d->kind == VB_F32_DOWN || (d->H % 2 == 0 && d->W % 2 == 0), "vb_f32_op: UP needs even output extent");
      vb::resample_f32_kernel<<<blocks(P * d->ca, 256), 256, 0, s>>>(*d, P * d->ca, d->kind == VB_F32_UP ? 1 : 0);
      break;
    case VB_F32_QKV:
      VB_REQUIRE(d->heads > 0 && d->parts >= 1 && d->parts <= 3 && d->head_dim > 0 && d->seg_div > 0, "vb_f32_op: bad QKV shape");
      VB_REQUIRE(d->ca == d->heads * d->parts * d->head_dim, "vb_f32_op: QKV channel count mismatch");
      VB_REQUIRE((d->parts < 2 || d->out2) && (d->parts < 3 || d->out3), "vb_f32_op: QKV output missing");
      vb::qkv_f32_kernel<<<blocks(P * d->heads * d->parts, 8), 256, 0, s>>>(*d);
      break;
    case VB_F32_PRECOND_IN:
      VB_REQUIRE(d->ca == (d->b ? 7 : 4), "vb_f32_op: PRECOND_IN writes 4 (or 7 with conditioning) channels");
      vb::precond_in_f32_kernel<<<blocks(P, 256), 256, 0, s>>>(*d);
      break;
    default:
      VB_REQUIRE(false, "vb_f32_op: unknown kind %d", d->kind);
  }
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_f32_attn(const float* q, const float* k, const float* v, float* y, int32_t B, int32_t heads, int32_t sq,
                           int32_t sk, int32_t head_dim, int32_t zero_keys, void* stream) {
  VB_REQUIRE(q && k && v && y, "vb_f32_attn: null tensor");
  VB_REQUIRE(B > 0 && heads > 0 && sq > 0 && sk > 0 && zero_keys >= 0, "vb_f32_attn: empty problem");
  VB_REQUIRE(head_dim == 64 || head_dim == 32, "vb_f32_attn: head_dim must be 32 or 64");
  const long long total = static_cast<long long>(B) * heads * sq;
  const unsigned grid = static_cast<unsigned>((total + 7) / 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (head_dim == 64) vb::attn_f32_kernel<64><<<grid, 256, 0, s>>>(q, k, v, y, heads, sq, sk, zero_keys, total);
  else vb::attn_f32_kernel<32><<<grid, 256, 0, s>>>(q, k, v, y, heads, sq, sk, zero_keys, total);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}
