// Fused elementwise passes over the 16-bit NHWC residual stream (fp32 math) — HBM-bound, 16-byte
// vector accesses, a power-of-two lane group per pixel so the per-pixel channel reductions are warp shuffles.
// Reference ops: normalize(dim=1) pixel-norm (training/models.py:171, 37-42), resample up/down
// (:48-61), mp_silu (:66-67), mp_cat (:78-84), MPFourier + embedding linears (:96-101, 388-391,
// 175), EDM preconditioning (NVPrecond.forward), Heun/guidance update (generate_images.py:62,
// 93-114) and the uint8 pixel codec (training/encoders.py:58-62).
#include <algorithm>

#include "common.h"
#include "ptx.cuh"

namespace vb {
namespace {

constexpr int kEwThreads = 256;
constexpr int kMaxVpl = 4;   // 16-byte vectors per lane -> up to 32*4*8 = 1024 channels per pixel

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Sum over the `lpp` (power of two) consecutive lanes that share a pixel.
__device__ __forceinline__ float group_sum(float v, int lpp) {
  for (int o = lpp >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
struct Vec8 {
  float v[8];
};
__device__ __forceinline__ Vec8 ld8(const op_t* p) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  Vec8 r;
  const float2 a = unpack_op2(q.x), b = unpack_op2(q.y), c = unpack_op2(q.z), d = unpack_op2(q.w);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
  return r;
}
__device__ __forceinline__ void st8(op_t* p, const Vec8& r) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_op2(r.v[0], r.v[1]), pack_op2(r.v[2], r.v[3]), pack_op2(r.v[4], r.v[5]),
                                            pack_op2(r.v[6], r.v[7]));
}
__device__ __forceinline__ uint4 pack8_silu(const Vec8& r, float scale) {
  return make_uint4(pack_op2(mp_silu_f(r.v[0] * scale), mp_silu_f(r.v[1] * scale)), pack_op2(mp_silu_f(r.v[2] * scale), mp_silu_f(r.v[3] * scale)),
                 pack_op2(mp_silu_f(r.v[4] * scale), mp_silu_f(r.v[5] * scale)), pack_op2(mp_silu_f(r.v[6] * scale), mp_silu_f(r.v[7] * scale)));
}
__device__ __forceinline__ void st8_silu(op_t* p, const Vec8& r, float scale) {
  *reinterpret_cast<uint4*>(p) = pack8_silu(r, scale);
}

// ---- PIXNORM / DOWN_PIXNORM: `lpp` lanes per output pixel (32/lpp pixels per warp) ------------------
template <bool DOWN>
__global__ void __launch_bounds__(kEwThreads) pixnorm_kernel(const op_t* __restrict__ a, op_t* __restrict__ out,
                                                             op_t* __restrict__ out_silu, long long pixels, int H, int W,
                                                             int C, int lpp) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int ppw = 32 / lpp;                         // pixels per warp
  const long long warp_id = (static_cast<long long>(blockIdx.x) * kEwThreads + threadIdx.x) >> 5;
  const long long pix = warp_id * ppw + lane / lpp;
  const int gl = lane % lpp;                        // lane within the pixel group
  const int nvec = C >> 3;
  const bool live = pix < pixels;
  Vec8 v[kMaxVpl];
  float ss = 0.f;
  if (live) {
    if (!DOWN) {
      const op_t* src = a + pix * C;
#pragma unroll
      for (int j = 0; j < kMaxVpl; ++j) {
        const int i = gl + j * lpp;
        if (i < nvec) {
          v[j] = ld8(src + i * 8);
#pragma unroll
          for (int e = 0; e < 8; ++e) ss += v[j].v[e] * v[j].v[e];
        }
      }
    } else {
      // output pixel (n, y, x) <- mean of the 2x2 input patch at (2y, 2x) of a [2H][2W] image
      const int x = static_cast<int>(pix % W);
      const long long t = pix / W;
      const int y = static_cast<int>(t % H);
      const long long n = t / H;
      const long long p00 = (n * 2 * H + 2 * y) * (2 * W) + 2 * x;
      const op_t* s0 = a + p00 * C;
      const op_t* s1 = a + (p00 + 1) * C;
      const op_t* s2 = a + (p00 + 2 * W) * C;
      const op_t* s3 = a + (p00 + 2 * W + 1) * C;
#pragma unroll
      for (int j = 0; j < kMaxVpl; ++j) {
        const int i = gl + j * lpp;
        if (i < nvec) {
          const Vec8 p = ld8(s0 + i * 8), q = ld8(s1 + i * 8), r = ld8(s2 + i * 8), s = ld8(s3 + i * 8);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            v[j].v[e] = 0.25f * (p.v[e] + q.v[e] + r.v[e] + s.v[e]);
            ss += v[j].v[e] * v[j].v[e];
          }
        }
      }
    }
  }
  ss = group_sum(ss, lpp);
  if (!live) return;
  const float inv = 1.0f / (1e-4f + sqrtf(ss) * rsqrtf(static_cast<float>(C)));
#pragma unroll
  for (int j = 0; j < kMaxVpl; ++j) {
    const int i = gl + j * lpp;
    if (i < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[j].v[e] *= inv;
      if (out) st8(out + pix * C + i * 8, v[j]);
      if (out_silu) st8_silu(out_silu + pix * C + i * 8, v[j], 1.0f);
    }
  }
}

// ---- UP: `lpp` lanes per INPUT pixel, each writes the 2x2 output patch -------------------------------
__global__ void __launch_bounds__(kEwThreads) up_kernel(const op_t* __restrict__ a, op_t* __restrict__ out,
                                                        op_t* __restrict__ out_silu, long long in_pixels, int Hi, int Wi,
                                                        int C, int lpp) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int ppw = 32 / lpp;
  const long long warp_id = (static_cast<long long>(blockIdx.x) * kEwThreads + threadIdx.x) >> 5;
  const long long pix = warp_id * ppw + lane / lpp;
  if (pix >= in_pixels) return;
  const int gl = lane % lpp;
  const int nvec = C >> 3;
  const int x = static_cast<int>(pix % Wi);
  const long long t = pix / Wi;
  const int y = static_cast<int>(t % Hi);
  const long long n = t / Hi;
  const long long o00 = (n * 2 * Hi + 2 * y) * (2 * Wi) + 2 * x;
  const long long offs[4] = {o00, o00 + 1, o00 + 2 * Wi, o00 + 2 * Wi + 1};
  for (int i = gl; i < nvec; i += lpp) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(a + pix * C + i * 8));
    const Vec8 v = ld8(a + pix * C + i * 8);
    const uint4 sl = pack8_silu(v, 1.0f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (out) *reinterpret_cast<uint4*>(out + offs[k] * C + i * 8) = raw;
      if (out_silu) *reinterpret_cast<uint4*>(out_silu + offs[k] * C + i * 8) = sl;
    }
  }
}

// ---- CAT / SILU: `lpp` lanes per pixel ------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads) cat_kernel(const op_t* __restrict__ a, const op_t* __restrict__ b,
                                                         op_t* __restrict__ out, op_t* __restrict__ out_silu,
                                                         long long pixels, int ca, int cb, float wa, float wb, int lpp) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int ppw = 32 / lpp;
  const long long warp_id = (static_cast<long long>(blockIdx.x) * kEwThreads + threadIdx.x) >> 5;
  const long long pix = warp_id * ppw + lane / lpp;
  if (pix >= pixels) return;
  const int gl = lane % lpp;
  const int na = ca >> 3, nb = cb >> 3;
  const int C = ca + cb;
  for (int i = gl; i < na + nb; i += lpp) {
    Vec8 v = i < na ? ld8(a + pix * ca + i * 8) : ld8(b + pix * cb + (i - na) * 8);
    const float w = i < na ? wa : wb;
#pragma unroll
    for (int e = 0; e < 8; ++e) v.v[e] *= w;
    if (out) st8(out + pix * C + i * 8, v);
    if (out_silu) st8_silu(out_silu + pix * C + i * 8, v, 1.0f);
  }
}

// ---- embedding ------------------------------------------------------------------------
// One block per batch row: Fourier features -> emb_noise (+ emb_label, mp_sum) -> mp_silu.
__global__ void __launch_bounds__(256) emb_kernel(const vb_emb_desc d) {
  pdl_grid_sync();
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
  const int lane = threadIdx.x & 31;
  // grid.y slices the output rows; a warp owns 8 rows at a time and issues all their (coalesced) weight loads before the
  // first reduction, so the L2 latency is paid once per 8 rows instead of once per row
  const int rows_per_slice = (d.cemb + gridDim.y - 1) / gridDim.y;
  const int j_end = min(d.cemb, (static_cast<int>(blockIdx.y) + 1) * rows_per_slice);
  for (int j0 = blockIdx.y * rows_per_slice + (threadIdx.x >> 5) * 8; j0 < j_end; j0 += (blockDim.x >> 5) * 8) {
    float e[8], g[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int j = min(j0 + r, d.cemb - 1);
      const float* wn = d.w_noise + static_cast<size_t>(j) * d.cnoise;
      float a = 0.f;
      for (int c = lane; c < d.cnoise; c += 32) a += __ldg(wn + c) * s_in[c];
      e[r] = a;
      float b2 = 0.f;
      if (d.w_label != nullptr) {
        const float* wl = d.w_label + static_cast<size_t>(j) * d.label_dim;
        for (int c = lane; c < d.label_dim; c += 32) b2 += __ldg(wl + c) * s_geo[c];
      }
      g[r] = b2;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float ev = warp_sum(e[r]);
      if (d.w_label != nullptr) ev = (ev * (1.f - t) + warp_sum(g[r]) * t) * inv;
      // full-precision silu here: this vector modulates every block
      if (lane == 0 && j0 + r < j_end) d.emb[static_cast<size_t>(b) * d.cemb + j0 + r] = ev / (1.f + expf(-ev)) * (1.0f / 0.596f);
    }
  }
}

// mod[b][m] = Wmod[m] . emb[b] + 1 for every block's emb_linear at once (reference Block.forward, models.py:175): a
// small fp32 GEMM, C[B x mod_total] = E[B x cemb] W^T.  Register-tiled SIMT: block = 64 channels x 32 batch rows,
// thread = 2 channels x 4 batch rows, K in shared-memory slices of 32.  (v1 — one warp per channel with a shuffle
// reduction per batch row — took ~85 us for 11 904 channels; the modulation must stay fp32, so no tensor cores here.)
constexpr int kModBM = 64, kModBK = 32;
__global__ void __launch_bounds__(256) mod_kernel(const vb_emb_desc d) {
  pdl_grid_sync();
  __shared__ float s_w[kModBM][kModBK + 1];
  __shared__ __align__(16) float s_e[kModBK][32];
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  const int m0 = blockIdx.x * kModBM, b0 = blockIdx.y * 32;
  const int lr = tid >> 5, lc = tid & 31;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int k0 = 0; k0 < d.cemb; k0 += kModBK) {
#pragma unroll
    for (int i = 0; i < kModBM / 8; ++i) {
      const int r = lr + 8 * i;
      s_w[r][lc] = (m0 + r < d.mod_total && k0 + lc < d.cemb) ? __ldg(d.w_mod + static_cast<size_t>(m0 + r) * d.cemb + k0 + lc) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int bb = lr + 8 * i;
      s_e[lc][bb] = (b0 + bb < d.B && k0 + lc < d.cemb) ? d.emb[static_cast<size_t>(b0 + bb) * d.cemb + k0 + lc] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kModBK; ++k) {
      const float4 e = *reinterpret_cast<const float4*>(&s_e[k][tx * 4]);
      const float w0 = s_w[ty * 2][k], w1 = s_w[ty * 2 + 1][k];
      acc[0][0] = fmaf(w0, e.x, acc[0][0]);
      acc[0][1] = fmaf(w0, e.y, acc[0][1]);
      acc[0][2] = fmaf(w0, e.z, acc[0][2]);
      acc[0][3] = fmaf(w0, e.w, acc[0][3]);
      acc[1][0] = fmaf(w1, e.x, acc[1][0]);
      acc[1][1] = fmaf(w1, e.y, acc[1][1]);
      acc[1][2] = fmaf(w1, e.z, acc[1][2]);
      acc[1][3] = fmaf(w1, e.w, acc[1][3]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int b = b0 + tx * 4 + j;
    if (b >= d.B) continue;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int m = m0 + ty * 2 + i;
      if (m < d.mod_total) d.mod[static_cast<size_t>(b) * d.mod_total + m] = acc[i][j] + 1.0f;
    }
  }
}

// ---- preconditioning ------------------------------------------------------------------
__global__ void __launch_bounds__(256) precond_in_kernel(const vb_precond_in_desc d) {
  pdl_grid_sync();
  __shared__ uint4 stage[8][32][8];     // im2col fast path: per-warp transpose so that the 128-byte rows leave coalesced
  const long long pix_raw = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long hw = static_cast<long long>(d.R) * d.R;
  const long long total = hw * d.B;
  const bool fast = d.im2col && d.cpad == 64;
  if (pix_raw >= total && !fast) return;
  const bool valid = pix_raw < total;
  const long long pix = valid ? pix_raw : total - 1;      // (fast path: out-of-range lanes compute a dummy pixel, store nothing)
  const long long n = pix / hw, s = pix - n * hw;
  float c_in = 1.f;
  if (d.sigma != nullptr) {
    const float sg = d.sigma[d.sigma_n == 1 ? 0 : n * d.sigma_stride];
    c_in = rsqrtf(d.sigma_data * d.sigma_data + sg * sg);
  }
  const float* xp = d.x + n * d.img_stride;
  const float* cp = d.cond ? d.cond + n * 3 * hw : nullptr;
  const float* np = d.noise ? d.noise + n * 3 * hw : nullptr;
  const int nch = cp ? 7 : 4;           // [x*c_in (3), cond + noisy_sr*noise (3, SR only), ones]
  auto channel = [&](int c, long long at) -> float {
    if (c < 3) return xp[c * hw + at] * c_in;
    if (c == nch - 1) return 1.0f;     // bias-as-channel (training/models.py:394)
    return cp[(c - 3) * hw + at] + (np ? d.noisy_sr * np[(c - 3) * hw + at] : 0.f);
  };
  op_t* o = static_cast<op_t*>(d.out) + pix * d.cpad;
  if (!d.im2col) {
    float ch[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < nch; ++c) ch[c] = channel(c, s);
    uint4* o4 = reinterpret_cast<uint4*>(o);
    o4[0] = make_uint4(pack_op2(ch[0], ch[1]), pack_op2(ch[2], ch[3]), pack_op2(ch[4], ch[5]), pack_op2(ch[6], ch[7]));
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int i = 1; i < d.cpad / 8; ++i) o4[i] = z;
    return;
  }
  // im2col: channel ci*9 + tap holds input channel ci at the tap's neighbour (zero outside the image, as conv padding does)
  const int y = static_cast<int>(s / d.R), x = static_cast<int>(s - static_cast<long long>(y) * d.R);
  if (d.cpad == 64) {
    // First convs at K = 64 (4 channels -> 36, SR's 7 -> 63): every neighbour load issued up front, straight-line
    // packing.  (The generic loop below ran at 1.2 TB/s of output on the 256x256 inputs: latency-bound behind
    // per-element branches.)
    bool ok[9];
    long long off[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      ok[tap] = yy >= 0 && yy < d.R && xx >= 0 && xx < d.R;
      off[tap] = ok[tap] ? static_cast<long long>(yy) * d.R + xx : s;
    }
    float v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = 0.f;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float a = __ldg(xp + ci * hw + off[tap]);
        v[ci * 9 + tap] = ok[tap] ? a * c_in : 0.f;
        if (cp) {
          const float c = __ldg(cp + ci * hw + off[tap]);
          const float nz = np ? __ldg(np + ci * hw + off[tap]) : 0.f;
          v[(ci + 3) * 9 + tap] = ok[tap] ? c + d.noisy_sr * nz : 0.f;
        }
      }
    }
    if (cp) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) v[54 + tap] = ok[tap] ? 1.0f : 0.f;
    } else {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) v[27 + tap] = ok[tap] ? 1.0f : 0.f;
    }
    // A thread owns one pixel = one 128-byte output row; written directly, every store instruction of the warp would
    // touch 32 different rows (16 B each).  Transpose through shared memory (XOR-swizzled 16-byte slots, conflict-free
    // both ways) so that 8 lanes write one whole row and a store instruction covers 4 contiguous rows.
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      stage[wid][lane][i ^ (lane & 7)] = make_uint4(pack_op2(v[8 * i], v[8 * i + 1]), pack_op2(v[8 * i + 2], v[8 * i + 3]),
                                                    pack_op2(v[8 * i + 4], v[8 * i + 5]), pack_op2(v[8 * i + 6], v[8 * i + 7]));
    __syncwarp();
    const long long pix0 = pix_raw - lane;                 // first pixel of this warp
    uint4* out4 = reinterpret_cast<uint4*>(d.out);
    const int c = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = 4 * i + (lane >> 3);
      if (pix0 + row < total) out4[(pix0 + row) * 8 + c] = stage[wid][row][c ^ (row & 7)];
    }
    return;
  }
  for (int base = 0; base < d.cpad; base += 8) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = base + e;
      const int ci = k / 9, tap = k - ci * 9;
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      v[e] = (ci < nch && yy >= 0 && yy < d.R && xx >= 0 && xx < d.R) ? channel(ci, static_cast<long long>(yy) * d.R + xx) : 0.f;
    }
    *reinterpret_cast<uint4*>(o + base) = make_uint4(pack_op2(v[0], v[1]), pack_op2(v[2], v[3]), pack_op2(v[4], v[5]), pack_op2(v[6], v[7]));
  }
}

__global__ void __launch_bounds__(256) precond_out_kernel(const vb_precond_out_desc d) {
  pdl_grid_sync();
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
    if (d.x_out[0]) reinterpret_cast<float4*>(d.x_out[0])[i4] = xn;
    if (d.x_out[1]) reinterpret_cast<float4*>(d.x_out[1])[i4] = xn;
  }
  if (i4 < d.sigma_n) {      // the next call's noise level, per sample
    if (d.sigma_out[0]) d.sigma_out[0][i4] = d.sigma_next;
    if (d.sigma_out[1]) d.sigma_out[1][i4] = d.sigma_next;
  }
  if (i4 == 0) {   // tail (n not a multiple of 4)
    for (long long i = n4 << 2; i < d.n; ++i) {
      float dc = d.phase == 0 ? 0.f : d.d_cur[i], xn = d.phase == 0 ? 0.f : d.x_next[i];
      step(d.d_net[i], d.d_gnet ? d.d_gnet[i] : 0.f, d.x_hat[i], dc, xn);
      if (d.phase == 0) d.d_cur[i] = dc;
      d.x_next[i] = xn;
      if (d.x_out[0]) d.x_out[0][i] = xn;
      if (d.x_out[1]) d.x_out[1][i] = xn;
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
inline int lanes_per_pixel(int channels) {
  const int nvec = channels / 8;
  int lpp = 1;
  while (lpp < nvec && lpp < 32) lpp <<= 1;
  return lpp;
}

}  // namespace

int eltwise_launch(const vb_ew_desc* d, cudaStream_t s) {
  VB_REQUIRE(d != nullptr && d->a != nullptr, "vb_eltwise: null input");
  VB_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0, "vb_eltwise: empty extent");
  VB_REQUIRE(d->ca > 0 && d->ca % 8 == 0 && d->ca <= kMaxVpl * 256, "vb_eltwise: channels %d must be a multiple of 8, <= %d",
             d->ca, kMaxVpl * 256);
  VB_REQUIRE(d->out || d->out_silu, "vb_eltwise: no output");
  const long long pixels = static_cast<long long>(d->B) * d->H * d->W;
  const op_t* a = static_cast<const op_t*>(d->a);
  op_t* o = static_cast<op_t*>(d->out);
  op_t* os = static_cast<op_t*>(d->out_silu);
  switch (d->kind) {
    case VB_EW_PIXNORM:
    case VB_EW_DOWN_PIXNORM: {
      const int lpp = lanes_per_pixel(d->ca);
      const long long warps = (pixels + 32 / lpp - 1) / (32 / lpp);
      if (d->kind == VB_EW_PIXNORM)
        VB_CHECK_CUDA(launch_pdl(pixnorm_kernel<false>, dim3(warp_grid(warps)), dim3(kEwThreads), 0, s, a, o, os, pixels, d->H, d->W, d->ca, lpp));
      else
        VB_CHECK_CUDA(launch_pdl(pixnorm_kernel<true>, dim3(warp_grid(warps)), dim3(kEwThreads), 0, s, a, o, os, pixels, d->H, d->W, d->ca, lpp));
      break;
    }
    case VB_EW_UP: {
      VB_REQUIRE(d->H % 2 == 0 && d->W % 2 == 0, "vb_eltwise: UP needs even output extent");
      const long long in_pixels = pixels / 4;
      const int lpp = lanes_per_pixel(d->ca);
      const long long warps = (in_pixels + 32 / lpp - 1) / (32 / lpp);
      VB_CHECK_CUDA(launch_pdl(up_kernel, dim3(warp_grid(warps)), dim3(kEwThreads), 0, s, a, o, os, in_pixels, d->H / 2, d->W / 2, d->ca, lpp));
      break;
    }
    case VB_EW_CAT:
    case VB_EW_SILU: {
      const bool cat = d->kind == VB_EW_CAT;
      if (cat) VB_REQUIRE(d->b != nullptr && d->cb > 0 && d->cb % 8 == 0, "vb_eltwise: CAT needs b with cb %% 8 == 0");
      const int lpp = lanes_per_pixel(d->ca + (cat ? d->cb : 0));
      const long long warps = (pixels + 32 / lpp - 1) / (32 / lpp);
      VB_CHECK_CUDA(launch_pdl(cat_kernel, dim3(warp_grid(warps)), dim3(kEwThreads), 0, s, a,
                               cat ? static_cast<const op_t*>(d->b) : nullptr, o, os, pixels, d->ca, cat ? d->cb : 0,
                               cat ? d->wa : (d->wa != 0.f ? d->wa : 1.0f), cat ? d->wb : 1.0f, lpp));
      break;
    }
    default:
      VB_REQUIRE(false, "vb_eltwise: unknown kind %d", d->kind);
  }
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// The same GEMM for batches above 32: block = 64 channels x 64 batch rows, thread = 4 x 4 outputs, both operands K-major in
// shared memory (two 16-byte shared loads per 16 FMAs instead of three loads per 8), the next K slice prefetched into registers
// while the current one is multiplied.  Every output is still the sum over k = 0 .. cemb-1 in ascending order in one fp32
// accumulator, so the bits equal the 32-row kernel's — a plan gives the same modulation whatever batch it is built for
// (tests/test_gpu_parity.py::test_embed_modulation_bits_do_not_depend_on_the_batch).  profiles/r02_step_timeline.txt: the 32-row
// kernel took 112 us per launch at batch 128 against a 29 us fp32-FMA floor, 0.5 % of the step; profiles/r02_embed_micro.txt: A/B.
constexpr int kModWBN = 64, kModWLd = kModBM + 4;
__global__ void __launch_bounds__(256) mod_wide_kernel(const vb_emb_desc d) {
  pdl_grid_sync();
  __shared__ __align__(16) float s_w[kModBK][kModWLd];     // [k][channel]
  __shared__ __align__(16) float s_e[kModBK][kModWLd];     // [k][batch row]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * kModBM, b0 = blockIdx.y * kModWBN;
  // global -> shared: a thread moves two float4 (4 consecutive k) of each operand per slice: rows lr and lr + 32, k = 4 kq .. 4 kq + 3
  const int lr = tid >> 3, kq = tid & 7;
  const bool vec_ok = (d.cemb & 3) == 0;
  auto fetch = [&](const float* base, int row, int rows, int k0) -> float4 {
    const int k = k0 + kq * 4;
    if (row >= rows) return make_float4(0.f, 0.f, 0.f, 0.f);
    const float* p = base + static_cast<size_t>(row) * d.cemb + k;
    if (vec_ok && k + 3 < d.cemb) return __ldg(reinterpret_cast<const float4*>(p));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < d.cemb) v.x = __ldg(p);
    if (k + 1 < d.cemb) v.y = __ldg(p + 1);
    if (k + 2 < d.cemb) v.z = __ldg(p + 2);
    if (k + 3 < d.cemb) v.w = __ldg(p + 3);
    return v;
  };
  auto stash = [&](float (*dst)[kModWLd], int row, const float4& v) {
    dst[kq * 4 + 0][row] = v.x;
    dst[kq * 4 + 1][row] = v.y;
    dst[kq * 4 + 2][row] = v.z;
    dst[kq * 4 + 3][row] = v.w;
  };
  float4 pw0 = fetch(d.w_mod, m0 + lr, d.mod_total, 0), pw1 = fetch(d.w_mod, m0 + lr + 32, d.mod_total, 0);
  float4 pe0 = fetch(d.emb, b0 + lr, d.B, 0), pe1 = fetch(d.emb, b0 + lr + 32, d.B, 0);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < d.cemb; k0 += kModBK) {
    stash(s_w, lr, pw0);
    stash(s_w, lr + 32, pw1);
    stash(s_e, lr, pe0);
    stash(s_e, lr + 32, pe1);
    __syncthreads();
    if (k0 + kModBK < d.cemb) {
      pw0 = fetch(d.w_mod, m0 + lr, d.mod_total, k0 + kModBK);
      pw1 = fetch(d.w_mod, m0 + lr + 32, d.mod_total, k0 + kModBK);
      pe0 = fetch(d.emb, b0 + lr, d.B, k0 + kModBK);
      pe1 = fetch(d.emb, b0 + lr + 32, d.B, k0 + kModBK);
    }
#pragma unroll
    for (int k = 0; k < kModBK; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(&s_w[k][ty * 4]);
      const float4 e = *reinterpret_cast<const float4*>(&s_e[k][tx * 4]);
      const float wv[4] = {w.x, w.y, w.z, w.w}, ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wv[i], ev[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int b = b0 + tx * 4 + j;
    if (b >= d.B) continue;
    float* o = d.mod + static_cast<size_t>(b) * d.mod_total + m0 + ty * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (m0 + ty * 4 + i < d.mod_total) o[i] = acc[i][j] + 1.0f;
  }
}

int embed_launch(const vb_emb_desc* d, cudaStream_t s) {
  VB_REQUIRE(d != nullptr && d->sigma && d->freqs && d->phases && d->w_noise && d->emb, "vb_embed: null argument");
  VB_REQUIRE(d->B > 0 && d->cnoise > 0 && d->cemb > 0, "vb_embed: empty problem");
  VB_REQUIRE(d->mod_total == 0 || (d->w_mod && d->mod), "vb_embed: w_mod/mod missing");
  const size_t smem = sizeof(float) * (d->cnoise + (d->w_label ? d->label_dim : 0));
  VB_CHECK_CUDA(launch_pdl(emb_kernel, dim3(d->B, 4), dim3(256), smem, s, *d));
  VB_CHECK_CUDA(cudaGetLastError());
  if (d->mod_total > 0) {
    static const bool wide_off = getenv("VB_MOD_WIDE") != nullptr && atoi(getenv("VB_MOD_WIDE")) == 0;      // A/B testing
    // the 64 x 64 tiles pay once they fill the chip (vivid-base at batch 128: 372 CTAs, vb_embed 147 -> 103 us; the SR UNet's 3 712
    // channels give 116 and run 44 -> 57 us, so they stay with the 32-row kernel); the choice never changes a bit
    const dim3 wide((d->mod_total + kModBM - 1) / kModBM, (d->B + kModWBN - 1) / kModWBN);
    if (d->B > 32 && !wide_off && static_cast<int>(wide.x * wide.y) >= num_sms()) {
      VB_CHECK_CUDA(launch_pdl(mod_wide_kernel, wide, dim3(256), 0, s, *d));
    } else {
      const dim3 grid((d->mod_total + kModBM - 1) / kModBM, (d->B + 31) / 32);
      VB_CHECK_CUDA(launch_pdl(mod_kernel, grid, dim3(256), 0, s, *d));
    }
    VB_CHECK_CUDA(cudaGetLastError());
  }
  return VB_OK;
}

int precond_in_launch(const vb_precond_in_desc* d, cudaStream_t s) {
  VB_REQUIRE(d != nullptr && d->x && d->out, "vb_precond_in: null tensor");
  VB_REQUIRE(d->B > 0 && d->R > 0 && d->cpad >= 8 && d->cpad % 8 == 0, "vb_precond_in: bad extent");
  VB_REQUIRE(!d->im2col || d->cpad >= 9 * (d->cond ? 7 : 4), "vb_precond_in: im2col needs cpad >= 9 * channels");
  const long long pixels = static_cast<long long>(d->B) * d->R * d->R;
  VB_CHECK_CUDA(launch_pdl(precond_in_kernel, dim3(static_cast<unsigned>((pixels + 255) / 256)), dim3(256), 0, s, *d));
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

int precond_out_launch(const vb_precond_out_desc* d, cudaStream_t s) {
  VB_REQUIRE(d != nullptr && d->x && d->f && d->sigma && d->d_out, "vb_precond_out: null tensor");
  VB_REQUIRE(d->B > 0 && d->R > 0 && d->ldf >= 4 && d->ldf % 4 == 0, "vb_precond_out: bad extent");
  const long long pixels = static_cast<long long>(d->B) * d->R * d->R;
  VB_CHECK_CUDA(launch_pdl(precond_out_kernel, dim3(static_cast<unsigned>((pixels + 255) / 256)), dim3(256), 0, s, *d));
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

// Uncertainty head u(sigma) (training/models.py:746-747, dual :686-688): a 1-output magnitude-preserving linear over the
// Fourier features of c_noise = ln(sigma)/4.  One block per sample; the weight normalisation (||w||, fp32) is folded in.
__global__ void __launch_bounds__(128) logvar_kernel(const float* __restrict__ sigma, int sigma_stride, const float* __restrict__ w,
                                                     const float* __restrict__ freqs, const float* __restrict__ phases, int C,
                                                     float* __restrict__ out) {
  __shared__ float red[2][4];
  const float c_noise = logf(sigma[static_cast<long long>(blockIdx.x) * sigma_stride]) * 0.25f;
  float dot = 0.f, sq = 0.f;
  for (int c = threadIdx.x; c < C; c += 128) {
    const float wc = w[c];
    dot = fmaf(wc, cosf(fmaf(c_noise, freqs[c], phases[c])) * 1.4142135623730951f, dot);
    sq = fmaf(wc, wc, sq);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = dot; red[1][threadIdx.x >> 5] = sq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    dot = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    sq = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    const float rc = rsqrtf(static_cast<float>(C));
    out[blockIdx.x] = dot * rc / (1e-4f + sqrtf(sq) * rc);
  }
}

int heun_launch(const vb_heun_desc* d, cudaStream_t s) {
  VB_REQUIRE(d != nullptr && d->d_net && d->x_hat && d->d_cur && d->x_next, "vb_heun: null tensor");
  VB_REQUIRE(d->n > 0 && (d->phase == 0 || d->phase == 1), "vb_heun: bad n/phase");
  const long long n4 = (d->n >> 2) > 0 ? (d->n >> 2) : 1;
  VB_REQUIRE(d->sigma_n >= 0 && d->sigma_n <= n4, "vb_heun: bad sigma_n");
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
extern "C" int vb_logvar(const float* sigma, int32_t n, int32_t sigma_stride, const float* weight, const float* freqs,
                         const float* phases, int32_t channels, float* out, void* stream) {
  VB_REQUIRE(sigma && weight && freqs && phases && out, "vb_logvar: null tensor");
  VB_REQUIRE(n > 0 && channels > 0 && sigma_stride >= 0, "vb_logvar: bad n/channels/stride");
  vb::logvar_kernel<<<static_cast<unsigned>(n), 128, 0, static_cast<cudaStream_t>(stream)>>>(sigma, sigma_stride, weight, freqs,
                                                                                             phases, channels, out);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}
