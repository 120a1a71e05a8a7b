// Implicit-GEMM convolution (3x3 same-pad / 1x1) for NHWC bf16 activations on the
// sm_100a tensor cores:  TMA (4-D tiled, shifted boxes, zero-filled halo) -> shared
// memory (SWIZZLE_128B) -> tcgen05.mma with fp32 accumulators in TMEM -> fused epilogue.
//
// Replaces MPConv.forward's F.conv2d (reference training/models.py:126) plus the
// pointwise ops Block.forward runs around it (:174-205) and the qkv normalise/split
// (:192-193, 283-297).  GEMM view: M = B*H*W pixels (tile = 128 pixels forming a
// bn x bh x bw patch), N = output channels (tile = block_n), K = taps * input channels
// (64 per pipeline stage).
//
// CTA = 8 warps, persistent over tiles (static round-robin schedule):
//   warp 0   TMA producer (one lane)            warp 1   MMA issuer (one lane)
//   warp 2   TMEM allocator                     warps 4-7 epilogue (one thread per pixel row)
// Pipelines: smem full/empty ring (TMA <-> MMA) and a double-buffered TMEM accumulator
// (MMA <-> epilogue), so the epilogue of tile i overlaps the main loop of tile i+1.
#include <algorithm>
#include <new>

#include "common.h"
#include "ptx.cuh"

namespace vb {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kAStageBytes = kBlockM * kBlockK * 2;   // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kThreads = 256;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;                       // columns between the two accumulator buffers
constexpr int kSmemBudget = 200 * 1024;

struct ConvKernelParams {
  int B, H, W;
  int bw, bh, bn;
  int tiles_x, tiles_y;
  int n_tiles, total_tiles;
  int taps, kc_a, kc_b;
  int block_n;
  int num_stages, stage_bytes, b_bytes;
  uint32_t idesc;
  int epi_mode, flags;
  const float* mod;
  int mod_stride;
  const float* res;
  int ld_res;
  float* out_f32;
  int ld_f32;
  __nv_bfloat16* out_bf16;
  int ld_bf16;
  __nv_bfloat16* out_silu;
  int ld_silu;
  float res_a, res_b, clip;
  int head_dim, parts, seg_div, heads;
  __nv_bfloat16* part0;
  __nv_bfloat16* part1;
  __nv_bfloat16* part2;
  int part_seq[3];
  int part_off[3];
  float norm_scale;   // 1/sqrt(head_dim)
};

struct TileCoord {
  int x0, y0, n0, col0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvKernelParams& p, int tile) {
  const int mt = tile / p.n_tiles;
  const int nt = tile - mt * p.n_tiles;
  const int tx = mt % p.tiles_x;
  const int t2 = mt / p.tiles_x;
  const int ty = t2 % p.tiles_y;
  const int tn = t2 / p.tiles_y;
  TileCoord t;
  t.x0 = tx * p.bw;
  t.y0 = ty * p.bh;
  t.n0 = tn * p.bn;
  t.col0 = nt * p.block_n;
  return t;
}

// One chunk of NC accumulator columns of one pixel row: modulation+mp_silu, mp_sum with the
// residual stream, clip, then up to three stores (fp32 stream, bf16 GEMM operand, bf16 mp_silu).
template <int NC>
__device__ __forceinline__ void epi_chunk_plain(const ConvKernelParams& p, uint32_t taddr, int col, size_t pix, int n,
                                                bool valid) {
  float v[NC];
  if (NC == 32)
    tmem_ld32(taddr, v);
  else
    tmem_ld16(taddr, v);
  tmem_ld_wait();
  if (!valid) return;
  if (p.flags & VB_F_MODSILU) {
    const float4* m = reinterpret_cast<const float4*>(p.mod + static_cast<size_t>(n) * p.mod_stride + col);
#pragma unroll
    for (int j = 0; j < NC / 4; ++j) {
      const float4 mm = __ldg(m + j);
      v[4 * j + 0] = mp_silu_f(v[4 * j + 0] * mm.x);
      v[4 * j + 1] = mp_silu_f(v[4 * j + 1] * mm.y);
      v[4 * j + 2] = mp_silu_f(v[4 * j + 2] * mm.z);
      v[4 * j + 3] = mp_silu_f(v[4 * j + 3] * mm.w);
    }
  }
  if (p.flags & VB_F_RESIDUAL) {
    const float4* r = reinterpret_cast<const float4*>(p.res + pix * p.ld_res + col);
#pragma unroll
    for (int j = 0; j < NC / 4; ++j) {
      const float4 rr = __ldg(r + j);
      v[4 * j + 0] = rr.x * p.res_a + v[4 * j + 0] * p.res_b;
      v[4 * j + 1] = rr.y * p.res_a + v[4 * j + 1] * p.res_b;
      v[4 * j + 2] = rr.z * p.res_a + v[4 * j + 2] * p.res_b;
      v[4 * j + 3] = rr.w * p.res_a + v[4 * j + 3] * p.res_b;
    }
  }
  if (p.flags & VB_F_CLIP) {
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = fminf(fmaxf(v[j], -p.clip), p.clip);
  }
  if (p.out_f32 != nullptr) {
    float4* o = reinterpret_cast<float4*>(p.out_f32 + pix * p.ld_f32 + col);
#pragma unroll
    for (int j = 0; j < NC / 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  if (p.out_bf16 != nullptr) {
    uint4* o = reinterpret_cast<uint4*>(p.out_bf16 + pix * p.ld_bf16 + col);
#pragma unroll
    for (int j = 0; j < NC / 8; ++j)
      o[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                        pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
  }
  if (p.out_silu != nullptr) {
    uint4* o = reinterpret_cast<uint4*>(p.out_silu + pix * p.ld_silu + col);
#pragma unroll
    for (int j = 0; j < NC / 8; ++j)
      o[j] = make_uint4(pack_bf16x2(mp_silu_f(v[8 * j]), mp_silu_f(v[8 * j + 1])),
                        pack_bf16x2(mp_silu_f(v[8 * j + 2]), mp_silu_f(v[8 * j + 3])),
                        pack_bf16x2(mp_silu_f(v[8 * j + 4]), mp_silu_f(v[8 * j + 5])),
                        pack_bf16x2(mp_silu_f(v[8 * j + 6]), mp_silu_f(v[8 * j + 7])));
  }
}

// One (head, q|k|v) group of D accumulator columns of one token: normalise over D in fp32
// (reference normalize(dim=2), eps 1e-4) and scatter to the [B][heads][seq][D] destination.
template <int D>
__device__ __forceinline__ void epi_group_qkv(const ConvKernelParams& p, uint32_t taddr, int gcol, int n, int s,
                                              bool valid) {
  float v[D];
  tmem_ld32(taddr, v);
  if (D == 64) tmem_ld32(taddr + 32, v + 32);
  tmem_ld_wait();
  if (!valid) return;
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < D; ++j) ss += v[j] * v[j];
  const float inv = 1.0f / (1e-4f + sqrtf(ss) * p.norm_scale);
  const int gg = gcol / D;
  const int part = gg % p.parts;
  const int head = gg / p.parts;
  const int b = n / p.seg_div;
  const int seg = n - b * p.seg_div;
  __nv_bfloat16* base = part == 0 ? p.part0 : (part == 1 ? p.part1 : p.part2);
  const int seq = part == 0 ? p.part_seq[0] : (part == 1 ? p.part_seq[1] : p.part_seq[2]);
  const int off = part == 0 ? p.part_off[0] : (part == 1 ? p.part_off[1] : p.part_off[2]);
  const size_t tok = (static_cast<size_t>(b) * p.heads + head) * seq + off + seg * (p.H * p.W) + s;
  uint4* o = reinterpret_cast<uint4*>(base + tok * D);
#pragma unroll
  for (int j = 0; j < D / 8; ++j)
    o[j] = make_uint4(pack_bf16x2(v[8 * j] * inv, v[8 * j + 1] * inv), pack_bf16x2(v[8 * j + 2] * inv, v[8 * j + 3] * inv),
                      pack_bf16x2(v[8 * j + 4] * inv, v[8 * j + 5] * inv),
                      pack_bf16x2(v[8 * j + 6] * inv, v[8 * j + 7] * inv));
}

__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                 const __grid_constant__ CUtensorMap map_w, const __grid_constant__ ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full[2];
  __shared__ __align__(8) uint64_t tmem_empty[2];
  __shared__ uint32_t tmem_slot;

  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    if (p.kc_b > 0) tma_prefetch_desc(&map_a2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = kAStageBytes + p.b_bytes;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        int kcol = 0;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = p.taps == 9 ? tap / 3 - 1 : 0;
          const int dx = p.taps == 9 ? tap % 3 - 1 : 0;
          for (int kc = 0; kc < p.kc_a + p.kc_b; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + stage * p.stage_bytes;
            uint8_t* sb = sa + kAStageBytes;
            mbar_expect_tx(&full_bar[stage], tx_bytes);
            if (kc < p.kc_a)
              tma_load_4d(&map_a, &full_bar[stage], sa, kc * kBlockK, t.x0 + dx, t.y0 + dy, t.n0);
            else
              tma_load_4d(&map_a2, &full_bar[stage], sa, (kc - p.kc_a) * kBlockK, t.x0 + dx, t.y0 + dy, t.n0);
            tma_load_2d(&map_w, &full_bar[stage], sb, kcol, t.col0);
            kcol += kBlockK;
            if (++stage == p.num_stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const int k_blocks = p.taps * (p.kc_a + p.kc_b);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t bphase = static_cast<uint32_t>(it >> 1) & 1u;
        mbar_wait(&tmem_empty[buf], bphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kAccStride);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
          const uint32_t sb = sa + kAStageBytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t adesc = umma_desc_sw128(sa + k * 32);
            const uint64_t bdesc = umma_desc_sw128(sb + k * 32);
            umma_bf16_ss(d_tmem, adesc, bdesc, p.idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);   // smem slot reusable once these MMAs retire
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tmem_full[buf]);       // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int rx = row % p.bw;
    const int r2 = row / p.bw;
    const int ry = r2 % p.bh;
    const int rn = r2 / p.bh;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t bphase = static_cast<uint32_t>(it >> 1) & 1u;
      const TileCoord t = decode_tile(p, tile);
      const int n = t.n0 + rn;
      const bool valid = n < p.B;
      const int s = (t.y0 + ry) * p.W + t.x0 + rx;
      const size_t pix = static_cast<size_t>(n) * p.H * p.W + s;
      mbar_wait(&tmem_full[buf], bphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(buf * kAccStride);
      if (p.epi_mode == VB_EPI_PLAIN) {
        int c = 0;
        for (; c + 32 <= p.block_n; c += 32) epi_chunk_plain<32>(p, taddr + c, t.col0 + c, pix, n, valid);
        if (c < p.block_n) epi_chunk_plain<16>(p, taddr + c, t.col0 + c, pix, n, valid);
      } else if (p.head_dim == 64) {
        for (int c = 0; c < p.block_n; c += 64) epi_group_qkv<64>(p, taddr + c, t.col0 + c, n, s, valid);
      } else {
        for (int c = 0; c < p.block_n; c += 32) epi_group_qkv<32>(p, taddr + c, t.col0 + c, n, s, valid);
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

struct ConvLaunch {
  CUtensorMap map_a, map_a2, map_w;
  ConvKernelParams p;
  int grid;
  int smem_bytes;
  double flops;
};

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

int conv_prepare(const vb_conv_desc* d, ConvLaunch** out) {
  VB_REQUIRE(d != nullptr && out != nullptr, "vb_conv: null descriptor");
  VB_REQUIRE(d->x != nullptr && d->w != nullptr, "vb_conv: x and w are required");
  VB_REQUIRE(d->B > 0 && is_pow2(d->H) && is_pow2(d->W) && d->H == d->W, "vb_conv: H=W must be a power of two (got %dx%d)",
             d->H, d->W);
  VB_REQUIRE(d->cin_pad > 0 && d->cin_pad % 64 == 0 && d->cin2_pad % 64 == 0 && d->cin2_pad >= 0,
             "vb_conv: input channels must be padded to multiples of 64 (got %d,%d)", d->cin_pad, d->cin2_pad);
  VB_REQUIRE((d->cin2_pad > 0) == (d->x2 != nullptr), "vb_conv: x2 and cin2_pad must come together");
  VB_REQUIRE(d->taps == 1 || d->taps == 9, "vb_conv: taps must be 1 or 9");
  VB_REQUIRE(d->block_n >= 16 && d->block_n <= 256 && d->block_n % 16 == 0, "vb_conv: block_n %d not in 16..256 step 16",
             d->block_n);
  VB_REQUIRE(d->cout_pad > 0 && d->cout_pad % d->block_n == 0, "vb_conv: cout_pad %d not a multiple of block_n %d",
             d->cout_pad, d->block_n);
  VB_REQUIRE(d->epi_mode == VB_EPI_PLAIN || d->epi_mode == VB_EPI_QKVNORM, "vb_conv: unknown epilogue %d", d->epi_mode);

  ConvLaunch* l = new (std::nothrow) ConvLaunch();
  VB_REQUIRE(l != nullptr, "vb_conv: out of host memory");
  ConvKernelParams& p = l->p;
  memset(&p, 0, sizeof(p));
  p.B = d->B;
  p.H = d->H;
  p.W = d->W;
  p.bw = std::min(d->W, kBlockM);
  p.bh = std::min(d->H, kBlockM / p.bw);
  p.bn = kBlockM / (p.bw * p.bh);
  p.tiles_x = d->W / p.bw;
  p.tiles_y = d->H / p.bh;
  const int tiles_nb = (d->B + p.bn - 1) / p.bn;
  p.n_tiles = d->cout_pad / d->block_n;
  p.total_tiles = p.tiles_x * p.tiles_y * tiles_nb * p.n_tiles;
  p.taps = d->taps;
  p.kc_a = d->cin_pad / 64;
  p.kc_b = d->cin2_pad / 64;
  p.block_n = d->block_n;
  p.b_bytes = d->block_n * 128;
  p.stage_bytes = kAStageBytes + p.b_bytes;
  p.num_stages = std::max(2, std::min(kMaxStages, kSmemBudget / p.stage_bytes));
  p.idesc = umma_idesc_bf16(kBlockM, d->block_n);
  p.epi_mode = d->epi_mode;
  p.flags = d->flags;
  p.mod = d->mod;
  p.mod_stride = d->mod_stride;
  p.res = d->res;
  p.ld_res = d->ld_res;
  p.out_f32 = d->out_f32;
  p.ld_f32 = d->ld_f32;
  p.out_bf16 = static_cast<__nv_bfloat16*>(d->out_bf16);
  p.ld_bf16 = d->ld_bf16;
  p.out_silu = static_cast<__nv_bfloat16*>(d->out_silu);
  p.ld_silu = d->ld_silu;
  const float t = d->res_t;
  const float inv = 1.0f / sqrtf((1.f - t) * (1.f - t) + t * t);
  p.res_a = (1.f - t) * inv;
  p.res_b = t * inv;
  p.clip = d->clip;

  auto fail = [&](int code) {
    delete l;
    return code;
  };
#define VB_REQUIRE_L(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      vb::set_error(__VA_ARGS__);    \
      return fail(VB_ERR_INVALID);   \
    }                                \
  } while (0)

  if (d->epi_mode == VB_EPI_PLAIN) {
    if (d->flags & VB_F_MODSILU) VB_REQUIRE_L(d->mod != nullptr && d->mod_stride % 4 == 0, "vb_conv: MODSILU needs mod (stride %% 4)");
    if (d->flags & VB_F_RESIDUAL) VB_REQUIRE_L(d->res != nullptr && d->ld_res % 4 == 0, "vb_conv: RESIDUAL needs res");
    VB_REQUIRE_L(d->out_f32 || d->out_bf16 || d->out_silu, "vb_conv: no output tensor");
    VB_REQUIRE_L((!d->out_f32 || d->ld_f32 % 4 == 0) && (!d->out_bf16 || d->ld_bf16 % 8 == 0) &&
                     (!d->out_silu || d->ld_silu % 8 == 0),
                 "vb_conv: output leading dimensions must keep 16-byte alignment");
  } else {
    VB_REQUIRE_L(d->head_dim == 64 || d->head_dim == 32, "vb_conv: head_dim must be 32 or 64");
    VB_REQUIRE_L(d->parts == 2 || d->parts == 3, "vb_conv: parts must be 2 (kv) or 3 (qkv)");
    VB_REQUIRE_L(d->block_n % d->head_dim == 0 && d->cout_pad % (d->parts * d->head_dim) == 0,
                 "vb_conv: qkv layout does not tile (cout %d, block_n %d, D %d)", d->cout_pad, d->block_n, d->head_dim);
    VB_REQUIRE_L(d->seg_div >= 1 && d->B % d->seg_div == 0, "vb_conv: seg_div must divide B");
    for (int j = 0; j < d->parts; ++j) VB_REQUIRE_L(d->part_out[j] != nullptr, "vb_conv: part_out[%d] missing", j);
    p.head_dim = d->head_dim;
    p.parts = d->parts;
    p.seg_div = d->seg_div;
    p.heads = d->cout_pad / (d->parts * d->head_dim);
    p.part0 = static_cast<__nv_bfloat16*>(d->part_out[0]);
    p.part1 = static_cast<__nv_bfloat16*>(d->part_out[1]);
    p.part2 = static_cast<__nv_bfloat16*>(d->part_out[2]);
    for (int j = 0; j < 3; ++j) {
      p.part_seq[j] = d->part_seq[j];
      p.part_off[j] = d->part_off[j];
    }
    p.norm_scale = 1.0f / sqrtf(static_cast<float>(d->head_dim));
  }

  // Tensor maps.  Activations: {C, W, H, N} with a {64, bw, bh, bn} box; weights: {K, cout_pad} with a {64, block_n} box.
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->cin_pad), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                              static_cast<uint64_t>(d->B)};
    const uint64_t strides[3] = {static_cast<uint64_t>(d->cin_pad) * 2, static_cast<uint64_t>(d->cin_pad) * 2 * d->W,
                                 static_cast<uint64_t>(d->cin_pad) * 2 * d->W * d->H};
    const uint32_t box[4] = {64, static_cast<uint32_t>(p.bw), static_cast<uint32_t>(p.bh), static_cast<uint32_t>(p.bn)};
    int rc = encode_tmap_bf16(&l->map_a, d->x, 4, dims, strides, box);
    if (rc != VB_OK) return fail(rc);
  }
  if (d->cin2_pad > 0) {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->cin2_pad), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                              static_cast<uint64_t>(d->B)};
    const uint64_t strides[3] = {static_cast<uint64_t>(d->cin2_pad) * 2, static_cast<uint64_t>(d->cin2_pad) * 2 * d->W,
                                 static_cast<uint64_t>(d->cin2_pad) * 2 * d->W * d->H};
    const uint32_t box[4] = {64, static_cast<uint32_t>(p.bw), static_cast<uint32_t>(p.bh), static_cast<uint32_t>(p.bn)};
    int rc = encode_tmap_bf16(&l->map_a2, d->x2, 4, dims, strides, box);
    if (rc != VB_OK) return fail(rc);
  } else {
    l->map_a2 = l->map_a;
  }
  {
    const uint64_t ktot = static_cast<uint64_t>(d->taps) * (d->cin_pad + d->cin2_pad);
    const uint64_t dims[2] = {ktot, static_cast<uint64_t>(d->cout_pad)};
    const uint64_t strides[1] = {ktot * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(d->block_n)};
    int rc = encode_tmap_bf16(&l->map_w, d->w, 2, dims, strides, box);
    if (rc != VB_OK) return fail(rc);
  }
#undef VB_REQUIRE_L

  l->grid = std::min(p.total_tiles, num_sms());
  l->smem_bytes = p.num_stages * p.stage_bytes + 1024;
  l->flops = 2.0 * d->B * d->H * d->W * static_cast<double>(d->cout_pad) * d->taps * (d->cin_pad + d->cin2_pad);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget + 8 * 1024);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(conv_gemm_kernel) failed: %s", cudaGetErrorString(e));
      return fail(VB_ERR_CUDA);
    }
    attr_done = true;
  }
  *out = l;
  return VB_OK;
}

int conv_launch(const ConvLaunch* l, cudaStream_t s) {
  conv_gemm_kernel<<<l->grid, kThreads, l->smem_bytes, s>>>(l->map_a, l->map_a2, l->map_w, l->p);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

void conv_free(ConvLaunch* l) { delete l; }
double conv_flops(const ConvLaunch* l) { return l->flops; }

}  // namespace vb

extern "C" int vb_conv(const vb_conv_desc* d, void* stream) {
  vb::ConvLaunch* l = nullptr;
  int rc = vb::conv_prepare(d, &l);
  if (rc != VB_OK) return rc;
  rc = vb::conv_launch(l, static_cast<cudaStream_t>(stream));
  vb::conv_free(l);
  return rc;
}
