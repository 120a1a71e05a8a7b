// Implicit-GEMM convolution (3x3 same-pad / 1x1) for NHWC 16-bit activations on the sm_100a
// tensor cores:  TMA (4-D tiled, shifted boxes, zero-filled halo) -> shared memory
// (SWIZZLE_128B) -> tcgen05.mma with fp32 accumulators in TMEM -> fused epilogue -> swizzled
// shared-memory staging -> TMA stores.
//
// Replaces MPConv.forward's F.conv2d (reference training/models.py:126) plus the pointwise ops
// Block.forward runs around it (:171-205: pixel-norm of the NEXT block, mp_silu, emb modulation,
// mp_sum, clip) and the qkv normalise/split (:192-193, 283-297).  GEMM view: M = B*H*W pixels
// (tile = 128 pixels forming a bn x bh x bw patch), N = output channels (tile = block_n),
// K = taps * input channels (64 per pipeline stage; up to two channel-concatenated sources, which
// is how mp_cat, :78-84, is folded into the consumer instead of being materialised).
//
// CTA = 8 warps, persistent over tiles (static round-robin schedule):
//   warp 0   TMA producer (one lane)            warp 1   MMA issuer (one lane)
//   warp 2   TMEM allocator                     warps 4-7 epilogue (one thread per pixel row)
// Pipelines: smem full/empty ring (TMA <-> MMA), a double-buffered TMEM accumulator (MMA <->
// epilogue) so the epilogue of tile i overlaps the main loop of tile i+1, and inside the epilogue
// a residual-tile TMA ring plus an output staging ring drained by TMA stores.
//
// Why the staged epilogue: one thread owns one pixel row of the accumulator (TMEM lane), so direct
// global loads/stores touch 32 different cache lines per warp instruction; ncu showed L1TEX at 69 %
// and the tensor pipe at 9.5 % on the K=576 layers (profiles/r01_conv_epilogue_before.txt).  TMA
// moves whole swizzled 128-byte rows and never enters L1TEX.
#include <algorithm>
#include <new>

#include "common.h"
#include "ptx.cuh"

namespace vb {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kAStageBytes = kBlockM * kBlockK * 2;   // 16 KiB
constexpr int kChunkBytes = kBlockM * 128;            // one 64-column 16-bit sub-tile: 16 KiB
constexpr int kMaxStages = 8;
constexpr int kThreads = 256;
constexpr int kEpiThreads = 128;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;                       // columns between the two accumulator buffers
constexpr int kSmemMax = 227 * 1024 - 2048;           // dynamic smem we allow ourselves (barriers are static)
constexpr int kEpiBarrier = 1;                        // named barrier id of the 4 epilogue warps

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
  int res_mode;         // VB_RES_*
  int res_resident;     // residual chunks of a tile fit the 2-buffer ring and are loaded once per tile
  int nslots;           // staged output slots in use
  int out_kind[3];
  float out_scale[3];
  int stage_depth;      // staging ring depth in chunks (1 or 2)
  int res_off, stg_off; // byte offsets of the residual ring / staging ring inside dynamic smem
  float* out_f32;
  int ld_f32;
  float res_a, res_b, clip;
  float inv_sqrt_c;     // 1/sqrt(cout) for the pixel norms
  int head_dim, parts, seg_div, heads;
  op_t* part0;
  op_t* part1;
  op_t* part2;
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

// 32 columns [c, c+32) of one pixel row: accumulator -> (modulation, mp_silu) -> mp_sum with the residual -> clip.
// `rrow` points at this row's 128-byte slot of the swizzled residual sub-tile (64 columns), `half` selects which 32.
__device__ __forceinline__ void compute_v32(const ConvKernelParams& p, uint32_t taddr, int col, int n, bool valid,
                                            const uint8_t* rrow, int row, int half, float res_inv, float* v) {
  tmem_ld32(taddr, v);
  tmem_ld_wait();
  if (p.flags & VB_F_MODSILU) {
    const float4* m = reinterpret_cast<const float4*>(p.mod + static_cast<size_t>(valid ? n : 0) * p.mod_stride + col);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 mm = __ldg(m + j);
      v[4 * j + 0] = mp_silu_fast(v[4 * j + 0] * mm.x);
      v[4 * j + 1] = mp_silu_fast(v[4 * j + 1] * mm.y);
      v[4 * j + 2] = mp_silu_fast(v[4 * j + 2] * mm.z);
      v[4 * j + 3] = mp_silu_fast(v[4 * j + 3] * mm.w);
    }
  }
  if (p.res_mode != VB_RES_NONE) {
    const float ra = p.res_a * res_inv;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 q = *reinterpret_cast<const uint4*>(rrow + (((half * 4 + j) ^ (row & 7)) << 4));
      const float2 a = unpack_op2(q.x), b = unpack_op2(q.y), c = unpack_op2(q.z), d = unpack_op2(q.w);
      v[8 * j + 0] = a.x * ra + v[8 * j + 0] * p.res_b;
      v[8 * j + 1] = a.y * ra + v[8 * j + 1] * p.res_b;
      v[8 * j + 2] = b.x * ra + v[8 * j + 2] * p.res_b;
      v[8 * j + 3] = b.y * ra + v[8 * j + 3] * p.res_b;
      v[8 * j + 4] = c.x * ra + v[8 * j + 4] * p.res_b;
      v[8 * j + 5] = c.y * ra + v[8 * j + 5] * p.res_b;
      v[8 * j + 6] = d.x * ra + v[8 * j + 6] * p.res_b;
      v[8 * j + 7] = d.y * ra + v[8 * j + 7] * p.res_b;
    }
  }
  if (p.flags & VB_F_CLIP) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fminf(fmaxf(v[j], -p.clip), p.clip);
  }
}

// Sum of squares of this row's 64 residual columns (for VB_RES_PIXNORM).
__device__ __forceinline__ float res_sumsq64(const uint8_t* rrow, int row) {
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 q = *reinterpret_cast<const uint4*>(rrow + ((j ^ (row & 7)) << 4));
    const float2 a = unpack_op2(q.x), b = unpack_op2(q.y), c = unpack_op2(q.z), d = unpack_op2(q.w);
    ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
  }
  return ss;
}

// Transform 32 values for one output slot and write them (16-bit) into this row of the staging sub-tile.
template <int KIND>
__device__ __forceinline__ void stage_out32_k(float scale, float inv_v, const float* v, uint8_t* srow, int row, int half) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float t[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float x = v[8 * j + e];
      t[e] = KIND == VB_OUT_RAW ? x
             : KIND == VB_OUT_SILU ? mp_silu_fast(x * scale)
             : KIND == VB_OUT_NORM ? x * inv_v
                                   : mp_silu_fast(x * inv_v);
    }
    *reinterpret_cast<uint4*>(srow + (((half * 4 + j) ^ (row & 7)) << 4)) =
        make_uint4(pack_op2(t[0], t[1]), pack_op2(t[2], t[3]), pack_op2(t[4], t[5]), pack_op2(t[6], t[7]));
  }
}
__device__ __forceinline__ void stage_out32(int kind, float scale, float inv_v, const float* v, uint8_t* srow, int row,
                                            int half) {
  switch (kind) {
    case VB_OUT_RAW: stage_out32_k<VB_OUT_RAW>(scale, inv_v, v, srow, row, half); break;
    case VB_OUT_SILU: stage_out32_k<VB_OUT_SILU>(scale, inv_v, v, srow, row, half); break;
    case VB_OUT_NORM: stage_out32_k<VB_OUT_NORM>(scale, inv_v, v, srow, row, half); break;
    default: stage_out32_k<VB_OUT_NORM_SILU>(scale, inv_v, v, srow, row, half); break;
  }
}

// QKVNORM: one (head, q|k|v) group of D accumulator columns of one token: normalise over D in fp32
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
  op_t* base = part == 0 ? p.part0 : (part == 1 ? p.part1 : p.part2);
  const int seq = part == 0 ? p.part_seq[0] : (part == 1 ? p.part_seq[1] : p.part_seq[2]);
  const int off = part == 0 ? p.part_off[0] : (part == 1 ? p.part_off[1] : p.part_off[2]);
  const size_t tok = (static_cast<size_t>(b) * p.heads + head) * seq + off + seg * (p.H * p.W) + s;
  uint4* o = reinterpret_cast<uint4*>(base + tok * D);
#pragma unroll
  for (int j = 0; j < D / 8; ++j)
    o[j] = make_uint4(pack_op2(v[8 * j] * inv, v[8 * j + 1] * inv), pack_op2(v[8 * j + 2] * inv, v[8 * j + 3] * inv),
                      pack_op2(v[8 * j + 4] * inv, v[8 * j + 5] * inv), pack_op2(v[8 * j + 6] * inv, v[8 * j + 7] * inv));
}

// fp32 direct-store epilogue for narrow outputs (out_conv: 16 padded columns).
__device__ __forceinline__ void epi_f32_16(const ConvKernelParams& p, uint32_t taddr, int col, size_t pix, bool valid) {
  float v[16];
  tmem_ld16(taddr, v);
  tmem_ld_wait();
  if (!valid) return;
  float4* o = reinterpret_cast<float4*>(p.out_f32 + pix * p.ld_f32 + col);
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

struct OutMaps {
  CUtensorMap m[3];
};

__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                 const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_res,
                 const __grid_constant__ OutMaps map_out, const __grid_constant__ ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full[2];
  __shared__ __align__(8) uint64_t tmem_empty[2];
  __shared__ __align__(8) uint64_t res_full[2];
  __shared__ uint32_t tmem_slot;

  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    if (p.kc_b > 0) tma_prefetch_desc(&map_a2);
    if (p.res_mode != VB_RES_NONE) tma_prefetch_desc(&map_res);
    for (int s = 0; s < p.nslots; ++s) tma_prefetch_desc(&map_out.m[s]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], kEpiThreads);
      mbar_init(&res_full[b], 1);
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
            umma_f16_ss(d_tmem, adesc, bdesc, p.idesc, (kb | k) != 0 ? 1u : 0u);
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
    const bool leader = threadIdx.x == kThreads - kEpiThreads;      // first epilogue thread issues the TMA traffic
    const int rx = row % p.bw;
    const int r2 = row / p.bw;
    const int ry = r2 % p.bh;
    const int rn = r2 / p.bh;
    const int chunks = p.block_n >> 6;                              // 64-column sub-tiles (staged path only)
    const bool staged = p.epi_mode == VB_EPI_PLAIN && p.nslots > 0;
    const bool has_res = p.res_mode != VB_RES_NONE;
    bool needs_norm = false;
    for (int s = 0; s < p.nslots; ++s) needs_norm |= p.out_kind[s] >= VB_OUT_NORM;
    uint8_t* res_ring = smem + p.res_off;
    uint8_t* stg_ring = smem + p.stg_off;
    uint32_t res_items = 0;      // residual sub-tiles consumed so far (ring position / phase), streaming mode
    uint32_t res_tiles = 0;      // tiles seen (phase of the resident residual buffers)
    uint32_t stg_chunks = 0;     // staging chunks produced so far (ring position)
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t bphase = static_cast<uint32_t>(it >> 1) & 1u;
      const TileCoord t = decode_tile(p, tile);
      const int n = t.n0 + rn;
      const bool valid = n < p.B;
      const int s_img = (t.y0 + ry) * p.W + t.x0 + rx;
      const size_t pix = static_cast<size_t>(n) * p.H * p.W + s_img;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(buf * kAccStride);

      if (!staged) {
        mbar_wait(&tmem_full[buf], bphase);
        tc_fence_after();
        if (p.epi_mode == VB_EPI_QKVNORM) {
          if (p.head_dim == 64) {
            for (int c = 0; c < p.block_n; c += 64) epi_group_qkv<64>(p, taddr + c, t.col0 + c, n, s_img, valid);
          } else {
            for (int c = 0; c < p.block_n; c += 32) epi_group_qkv<32>(p, taddr + c, t.col0 + c, n, s_img, valid);
          }
        } else {
          for (int c = 0; c < p.block_n; c += 16) epi_f32_16(p, taddr + c, t.col0 + c, pix, valid);
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty[buf]);
        continue;
      }

      // ---- residual sub-tiles: resident (<= 2 chunks, loaded once per tile) or streamed through the 2-buffer ring
      const int n_pass = 1 + (needs_norm ? 1 : 0) + (p.res_mode == VB_RES_PIXNORM ? 1 : 0);
      const int items = chunks * n_pass;                             // streaming: (pass, chunk) items of this tile
      if (has_res && leader) {
        if (p.res_resident) {
          for (int c = 0; c < chunks; ++c) {
            mbar_expect_tx(&res_full[c], kChunkBytes);
            tma_load_4d(&map_res, &res_full[c], res_ring + c * kChunkBytes, t.col0 + c * 64, t.x0, t.y0, t.n0);
          }
        } else {
          for (int k = 0; k < 2 && k < items; ++k) {
            const uint32_t slot = (res_items + k) & 1u;
            mbar_expect_tx(&res_full[slot], kChunkBytes);
            tma_load_4d(&map_res, &res_full[slot], res_ring + slot * kChunkBytes, t.col0 + (k % chunks) * 64, t.x0, t.y0, t.n0);
          }
        }
      }
      int item = 0;
      // Makes residual chunk c of the current pass readable; returns this row's 128-byte slot.
      auto res_acquire = [&](int c) -> const uint8_t* {
        if (!has_res) return nullptr;
        if (p.res_resident) {
          if (item < chunks) mbar_wait(&res_full[c], res_tiles & 1u);       // first touch this tile
          return res_ring + c * kChunkBytes + row * 128;
        }
        const uint32_t k = res_items + item;
        mbar_wait(&res_full[k & 1u], (k >> 1) & 1u);
        return res_ring + (k & 1u) * kChunkBytes + row * 128;
      };
      // Streaming mode: everybody is done with the ring slot of the current item -> refill it with item+2.
      auto res_release = [&](bool already_synced) {
        if (has_res && !p.res_resident) {
          if (!already_synced) named_bar_sync(kEpiBarrier, kEpiThreads);
          if (leader && item + 2 < items) {
            const uint32_t slot = (res_items + item) & 1u;
            mbar_expect_tx(&res_full[slot], kChunkBytes);
            tma_load_4d(&map_res, &res_full[slot], res_ring + slot * kChunkBytes, t.col0 + ((item + 2) % chunks) * 64, t.x0,
                        t.y0, t.n0);
          }
        }
        ++item;
      };

      // ---- pass: pixel-norm statistic of the residual row (VB_RES_PIXNORM)
      float res_inv = 1.0f;
      if (p.res_mode == VB_RES_PIXNORM) {
        float ss = 0.f;
        for (int c = 0; c < chunks; ++c) {
          const uint8_t* rrow = res_acquire(c);
          ss += res_sumsq64(rrow, row);
          res_release(false);
        }
        res_inv = 1.0f / (1e-4f + sqrtf(ss) * p.inv_sqrt_c);
      }

      mbar_wait(&tmem_full[buf], bphase);
      tc_fence_after();

      // ---- pass: pixel-norm statistic of the result row (VB_OUT_NORM*)
      float inv_v = 1.0f;
      if (needs_norm) {
        float ss = 0.f;
        for (int c = 0; c < chunks; ++c) {
          const uint8_t* rrow = res_acquire(c);
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            float v[32];
            compute_v32(p, taddr + c * 64 + half * 32, t.col0 + c * 64 + half * 32, n, valid, rrow, row, half, res_inv, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) ss += v[j] * v[j];
          }
          res_release(false);
        }
        inv_v = 1.0f / (1e-4f + sqrtf(ss) * p.inv_sqrt_c);
      }

      // ---- pass: outputs -> staging ring -> TMA stores
      for (int c = 0; c < chunks; ++c) {
        const uint8_t* rrow = res_acquire(c);
        // the staging entry about to be overwritten must have been read out by its TMA store
        if (stg_chunks >= static_cast<uint32_t>(p.stage_depth)) {
          if (leader) {
            if (p.stage_depth == 2)
              bulk_wait_read<1>();
            else
              bulk_wait_read<0>();
          }
          named_bar_sync(kEpiBarrier, kEpiThreads);
        }
        uint8_t* stg = stg_ring + (stg_chunks % p.stage_depth) * p.nslots * kChunkBytes;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          float v[32];
          compute_v32(p, taddr + c * 64 + half * 32, t.col0 + c * 64 + half * 32, n, valid, rrow, row, half, res_inv, v);
          if (p.out_f32 != nullptr && valid) {
            float4* o = reinterpret_cast<float4*>(p.out_f32 + pix * p.ld_f32 + t.col0 + c * 64 + half * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          for (int s = 0; s < p.nslots; ++s)
            stage_out32(p.out_kind[s], p.out_scale[s], inv_v, v, stg + s * kChunkBytes + row * 128, row, half);
        }
        if (c == chunks - 1) {
          tc_fence_before();
          mbar_arrive(&tmem_empty[buf]);      // accumulator fully consumed: MMA may start the tile after next
        }
        fence_proxy_async();                  // generic-proxy smem writes -> visible to the TMA engine
        named_bar_sync(kEpiBarrier, kEpiThreads);
        if (leader) {
          for (int s = 0; s < p.nslots; ++s)
            tma_store_4d(&map_out.m[s], stg + s * kChunkBytes, t.col0 + c * 64, t.x0, t.y0, t.n0);
          bulk_commit();
        }
        res_release(true);
        ++stg_chunks;
      }
      res_items += has_res && !p.res_resident ? items : 0;
      ++res_tiles;
    }
    if (leader) bulk_wait_read<0>();          // staging smem must outlive the last TMA store's read
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
  CUtensorMap map_a, map_a2, map_w, map_res;
  OutMaps map_out;
  ConvKernelParams p;
  int grid;
  int smem_bytes;
  double flops;
};

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

static int encode_act_map(CUtensorMap* map, const void* base, int C, int W, int H, int B, int bw, int bh, int bn) {
  const uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(C) * 2 * W,
                               static_cast<uint64_t>(C) * 2 * W * H};
  const uint32_t box[4] = {64, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), static_cast<uint32_t>(bn)};
  return encode_tmap_16(map, base, 4, dims, strides, box);
}

int conv_prepare(const vb_conv_desc* d, ConvLaunch** out) {
  VB_REQUIRE(d != nullptr && out != nullptr, "vb_conv: null descriptor");
  VB_REQUIRE(d->x != nullptr && d->w != nullptr, "vb_conv: x and w are required");
  VB_REQUIRE(d->B > 0 && is_pow2(d->H) && is_pow2(d->W) && d->H == d->W && d->H >= 4,
             "vb_conv: H=W must be a power of two >= 4 (got %dx%d)", d->H, d->W);
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
  p.idesc = umma_idesc_op(kBlockM, d->block_n);
  p.epi_mode = d->epi_mode;
  p.flags = d->flags;
  p.mod = d->mod;
  p.mod_stride = d->mod_stride;
  p.out_f32 = d->out_f32;
  p.ld_f32 = d->ld_f32;
  const float t = d->res_t;
  const float inv = 1.0f / sqrtf((1.f - t) * (1.f - t) + t * t);
  p.res_a = (1.f - t) * inv;
  p.res_b = t * inv;
  p.clip = d->clip;
  p.inv_sqrt_c = 1.0f / sqrtf(static_cast<float>(d->cout_pad));

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

  int epi_bytes = 0;
  if (d->epi_mode == VB_EPI_PLAIN) {
    if (d->flags & VB_F_MODSILU) VB_REQUIRE_L(d->mod != nullptr && d->mod_stride % 4 == 0, "vb_conv: MODSILU needs mod (stride %% 4)");
    bool needs_norm = false;
    for (int s = 0; s < 3; ++s) {
      if (d->out[s] == nullptr || d->out_kind[s] == VB_OUT_NONE) continue;
      VB_REQUIRE_L(d->out_kind[s] >= VB_OUT_RAW && d->out_kind[s] <= VB_OUT_NORM_SILU, "vb_conv: bad out_kind[%d]", s);
      VB_REQUIRE_L(p.nslots == s, "vb_conv: output slots must be filled in order");
      p.out_kind[p.nslots] = d->out_kind[s];
      p.out_scale[p.nslots] = d->out_scale[s] != 0.f ? d->out_scale[s] : 1.0f;
      needs_norm |= d->out_kind[s] >= VB_OUT_NORM;
      ++p.nslots;
    }
    p.res_mode = d->res_mode;
    VB_REQUIRE_L(d->res_mode >= VB_RES_NONE && d->res_mode <= VB_RES_PIXNORM, "vb_conv: bad res_mode");
    VB_REQUIRE_L((d->res_mode != VB_RES_NONE) == (d->res != nullptr), "vb_conv: res and res_mode must come together");
    VB_REQUIRE_L(p.nslots > 0 || d->out_f32 != nullptr, "vb_conv: no output tensor");
    if (p.nslots > 0) {
      VB_REQUIRE_L(d->block_n % 64 == 0, "vb_conv: staged 16-bit outputs need block_n %% 64 == 0 (got %d)", d->block_n);
      if (d->out_f32) VB_REQUIRE_L(d->ld_f32 % 4 == 0, "vb_conv: ld_f32 must keep 16-byte alignment");
    } else {
      VB_REQUIRE_L(d->res_mode == VB_RES_NONE && !(d->flags & (VB_F_MODSILU | VB_F_CLIP)) && d->ld_f32 % 4 == 0,
                   "vb_conv: the fp32-only epilogue is plain (no residual/modulation/clip)");
    }
    if (needs_norm || d->res_mode == VB_RES_PIXNORM)
      VB_REQUIRE_L(p.n_tiles == 1, "vb_conv: pixel-norm fusion needs the whole channel extent in one tile (cout_pad %d, block_n %d)",
                   d->cout_pad, d->block_n);
    const int chunks = d->block_n / 64;
    p.res_resident = chunks <= 2 ? 1 : 0;
    if (p.nslots > 0) {
      const int res_bytes = d->res_mode != VB_RES_NONE ? 2 * kChunkBytes : 0;
      p.stage_depth = 2;
      int stages = (kSmemMax - res_bytes - 2 * p.nslots * kChunkBytes) / p.stage_bytes;
      if (stages < 3) {
        p.stage_depth = 1;
        stages = (kSmemMax - res_bytes - p.nslots * kChunkBytes) / p.stage_bytes;
      }
      VB_REQUIRE_L(stages >= 2, "vb_conv: shared memory budget exceeded");
      epi_bytes = res_bytes + p.stage_depth * p.nslots * kChunkBytes;
    }
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
    p.part0 = static_cast<op_t*>(d->part_out[0]);
    p.part1 = static_cast<op_t*>(d->part_out[1]);
    p.part2 = static_cast<op_t*>(d->part_out[2]);
    for (int j = 0; j < 3; ++j) {
      p.part_seq[j] = d->part_seq[j];
      p.part_off[j] = d->part_off[j];
    }
    p.norm_scale = 1.0f / sqrtf(static_cast<float>(d->head_dim));
  }
  p.num_stages = std::max(2, std::min(kMaxStages, (kSmemMax - epi_bytes) / p.stage_bytes));
  p.res_off = p.num_stages * p.stage_bytes;
  p.stg_off = p.res_off + (p.res_mode != VB_RES_NONE ? 2 * kChunkBytes : 0);

  // Tensor maps.  Activations / residual / outputs: {C, W, H, N} with a {64, bw, bh, bn} box;
  // weights: {K, cout_pad} with a {64, block_n} box.
  int rc = encode_act_map(&l->map_a, d->x, d->cin_pad, d->W, d->H, d->B, p.bw, p.bh, p.bn);
  if (rc != VB_OK) return fail(rc);
  l->map_a2 = l->map_a;
  if (d->cin2_pad > 0) {
    rc = encode_act_map(&l->map_a2, d->x2, d->cin2_pad, d->W, d->H, d->B, p.bw, p.bh, p.bn);
    if (rc != VB_OK) return fail(rc);
  }
  l->map_res = l->map_a;
  if (p.res_mode != VB_RES_NONE) {
    rc = encode_act_map(&l->map_res, d->res, d->cout_pad, d->W, d->H, d->B, p.bw, p.bh, p.bn);
    if (rc != VB_OK) return fail(rc);
  }
  for (int s = 0; s < 3; ++s) l->map_out.m[s] = l->map_a;
  for (int s = 0; s < p.nslots; ++s) {
    rc = encode_act_map(&l->map_out.m[s], d->out[s], d->cout_pad, d->W, d->H, d->B, p.bw, p.bh, p.bn);
    if (rc != VB_OK) return fail(rc);
  }
  {
    const uint64_t ktot = static_cast<uint64_t>(d->taps) * (d->cin_pad + d->cin2_pad);
    const uint64_t dims[2] = {ktot, static_cast<uint64_t>(d->cout_pad)};
    const uint64_t strides[1] = {ktot * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(d->block_n)};
    rc = encode_tmap_16(&l->map_w, d->w, 2, dims, strides, box);
    if (rc != VB_OK) return fail(rc);
  }
#undef VB_REQUIRE_L

  l->grid = std::min(p.total_tiles, num_sms());
  l->smem_bytes = p.stg_off + p.stage_depth * p.nslots * kChunkBytes + 1024;
  l->flops = 2.0 * d->B * d->H * d->W * static_cast<double>(d->cout_pad) * d->taps * (d->cin_pad + d->cin2_pad);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax + 1024);
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
  conv_gemm_kernel<<<l->grid, kThreads, l->smem_bytes, s>>>(l->map_a, l->map_a2, l->map_w, l->map_res, l->map_out, l->p);
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
