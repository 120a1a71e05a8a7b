// Fused cosine attention on the sm_100a tensor cores: S = Q K^T and O = P V are tcgen05.mma with fp32 accumulators in
// TMEM; Q/K/V tiles arrive by TMA (SWIZZLE_128B); the softmax runs on 8 warps between the two products.
//
// Replaces einsum -> softmax -> einsum (snapshot experiments/code/training/models.py:190-191, 274-280) and
// F.scaled_dot_product_attention (training/models.py:198, 305) for head_dim 64 and sequence lengths that are multiples
// of 128 (every vivid-base / vivid-uncond attention above the 8x8 level); attention.cu (mma.sync) keeps the rest.
// q, k, v are pixel-normalised by the qkv GEMM epilogue, so |q.k|/sqrt(D) <= 8 and p = exp(logit) needs no running
// maximum and no rescaling of O (see attention.cu).
//
// One CTA = 12 warps, persistent over (batch, head, 256-query block) items = two 128-row query tiles that share every
// K/V block:
//   warps 0, 3  TMA producers: the two Q tiles once per item + the K ring (warp 0), the V ring (warp 3); 128 keys x 64 per tile
//   warp 1  MMA issuer of the logits S = Q K^T (runs up to three 128 x 128 tiles ahead: S ring of three TMEM buffers)
//   warp 2  TMEM allocator, then MMA issuer of O += P V (two accumulators of 64 columns, one per query tile)
//   warps 4-11 softmax + output: two groups of four warps on alternate key blocks; thread <-> one query row (TMEM lane):
//           tcgen05.ld -> 2^x on packed 16-bit pairs (one MUFU per two logits) -> P written to shared memory in the
//           K-major SWIZZLE_128B operand layout (A of the second product); row sums accumulate from the packed pairs.
// V is consumed as it lands ([key][d], d contiguous): that is the canonical MN-major SWIZZLE_128B B operand (instruction
// descriptor bit 16), so no transpose is needed anywhere.
#include <algorithm>

#include "common.h"
#include "ptx.cuh"

namespace vb {
namespace {

constexpr int kTQ = 128;                 // query rows per tile; an item is TWO tiles (256 rows) sharing every K/V block
constexpr int kTK = 128;                 // keys per block
constexpr int kHD = 64;                  // head dim
constexpr int kTile = kTQ * kHD * 2;     // 16 KiB: one 128 x 64 16-bit tile (Q, K, V, or half of P)
constexpr int kKvStages = 4;             // K ring and V ring, 16 KiB tiles each
constexpr int kAtThreads = 384;
constexpr int kGroupWarps = 4;           // softmax warps per query tile
constexpr int kOffQ = 0;                                   // Q tile A, Q tile B
constexpr int kOffK = 2 * kTile;                           // K ring
constexpr int kOffV = kOffK + kKvStages * kTile;           // V ring
constexpr int kOffP = kOffV + kKvStages * kTile;           // P of tile A, P of tile B: two 64-key sub-tiles each
constexpr int kAtSmem = kOffP + 2 * 2 * kTile + 1024;      // + alignment slack
constexpr int kColS = 0, kSBufs = 3, kColO = kSBufs * kTK, kAtCols = 512;     // S ring of 3 x 128 columns, O_A, O_B 64 each

struct AttnTcParams {
  int n_items, q_pairs, n_kv;        // items = B*heads*q_pairs; key blocks per item
  int heads, sq, sk;
  // tiles = 2: an item is two 128-row query tiles of one (batch, head) sharing every K/V block (sq % 256 == 0).
  // tiles = 1, small = 1 (sq == 64, the 8x8 level): an item is ONE 128-row tile holding the queries of TWO consecutive
  // (batch, head) pairs; its keys are the 2 sk rows of both (contiguous in [B][heads][sk][D]) and a 64-key sub-block only
  // counts for the rows of the pair it belongs to (block-diagonal logits: the other half of P is written as zeros).
  int tiles, small, n_bh;
  int out_d;                         // real head dim written out (32: q/k/v rows are zero-padded to 64 elements)
  float c1;                          // log2(e)/sqrt(D) still to be applied to the logits (1 if folded into q)
  float zero_keys;
  op_t* y;
  uint32_t idesc_s, idesc_o;
  int dbg;                           // VB_ATTN_DBG ablations (micro-benchmarks): 1 no logit scaling, 2 no exp, 4 no row sums,
                                     // 8 softmax handshakes only, 16 no P.V products, 32 no Q.K products
};

__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2" VB_WAIT_HINT ";\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins > VB_SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_2d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void commit_to(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t pin(uint32_t v) {
  uint32_t r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ uint32_t pk2(float lo, float hi) {
#ifdef VB_OP_BF16
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
#else
  __half2 h = __floats2half2_rn(lo, hi);
#endif
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t ex2_pk2(uint32_t x) {
  uint32_t r;
#ifdef VB_OP_BF16
  asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(r) : "r"(x));
#else
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(x));
#endif
  return r;
}
// 2^x of a packed fp16 pair on the FMA/ALU pipes (no MUFU): x = r + f with r = rint(x) (magic-number rounding: x + 1536 has
// an ulp of 1 in fp16, its low mantissa bits are r + 512), f in [-0.5, 0.5]; 2^f by a cubic (least-squares fit weighted for
// relative error; evaluated in fp16 its error is that of rounding the exact value: max 6.5e-4, mean 1.8e-4 relative over
// |x| <= 11.6 — the softmax's range, |q.k| log2(e)/sqrt(D) <= 8 log2(e)); 2^r is added into the exponent field with lane-wise
// integer arithmetic arranged so that no carry crosses the 16-bit lanes.  ex2.approx.f16x2 costs TWO MUFU issues on sm_100a
// and the softmax is MUFU-bound (profiles/r01_attn_tc.txt): every POLY-th pair takes this path instead.
__device__ __forceinline__ uint32_t ex2_poly_pk2(uint32_t x2) {
#ifdef VB_OP_BF16
  return ex2_pk2(x2);
#else
  const __half2 x = *reinterpret_cast<__half2*>(&x2);
  const __half2 magic = __float2half2_rn(1536.f);
  const __half2 y = __hadd2(x, magic);
  const __half2 r = __hsub2(y, magic);
  const __half2 f = __hsub2(x, r);
  __half2 pl = __hfma2(f, __float2half2_rn(0.05460186f), __float2half2_rn(0.24192564f));
  pl = __hfma2(pl, f, __float2half2_rn(0.69331678f));
  pl = __hfma2(pl, f, __float2half2_rn(1.0f));
  const uint32_t yb = *reinterpret_cast<const uint32_t*>(&y);
  const uint32_t t = ((yb & 0x03FF03FFu) - 0x01F001F0u) << 10;       // (r + 16) << 10 per lane, r in [-12, 12]
  return *reinterpret_cast<uint32_t*>(&pl) + t - 0x40004000u;
#endif
}
__device__ __forceinline__ uint32_t add_pk2(uint32_t a, uint32_t b) {
#ifdef VB_OP_BF16
  __nv_bfloat162 r = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
#else
  __half2 r = __hadd2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
#endif
  return *reinterpret_cast<uint32_t*>(&r);
}

template <bool SCALE, int POLY, bool SMALL>
__global__ void __launch_bounds__(kAtThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
               const __grid_constant__ CUtensorMap map_v, const __grid_constant__ AttnTcParams p) {
  constexpr int kTiles = SMALL ? 1 : 2;      // query tiles per item (see AttnTcParams)
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t b_q_full, b_q_empty;
  __shared__ __align__(8) uint64_t b_k_full[kKvStages], b_k_empty[kKvStages], b_v_full[kKvStages], b_v_empty[kKvStages];
  __shared__ __align__(8) uint64_t b_s_full[kSBufs], b_s_empty[kSBufs];
  __shared__ __align__(8) uint64_t b_p_full[2], b_p_empty[2], b_o_full[2], b_o_empty[2];
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&b_q_full, 1);
    mbar_init(&b_q_empty, 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(&b_k_full[s], 1);
      mbar_init(&b_k_empty[s], 1);
      mbar_init(&b_v_full[s], 1);
      mbar_init(&b_v_empty[s], 1);
    }
    for (int b = 0; b < kSBufs; ++b) {
      mbar_init(&b_s_full[b], 1);
      mbar_init(&b_s_empty[b], kGroupWarps);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&b_p_full[t], kGroupWarps);
      mbar_init(&b_p_empty[t], 1);
      mbar_init(&b_o_full[t], 1);
      mbar_init(&b_o_empty[t], kGroupWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_slot, kAtCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_grid_sync();      // set-up above overlaps the previous kernel's tail (launch_pdl)
  const uint32_t tmem = tmem_slot;
  const uint32_t sm = pin(smem_u32(smem));
  const uint32_t q_full = pin(smem_u32(&b_q_full)), q_empty = pin(smem_u32(&b_q_empty));
  const uint32_t k_full = pin(smem_u32(&b_k_full[0])), k_empty = pin(smem_u32(&b_k_empty[0]));
  const uint32_t v_full = pin(smem_u32(&b_v_full[0])), v_empty = pin(smem_u32(&b_v_empty[0]));
  const uint32_t s_full = pin(smem_u32(&b_s_full[0])), s_empty = pin(smem_u32(&b_s_empty[0]));
  const uint32_t p_full = pin(smem_u32(&b_p_full[0])), p_empty = pin(smem_u32(&b_p_empty[0]));
  const uint32_t o_full = pin(smem_u32(&b_o_full[0])), o_empty = pin(smem_u32(&b_o_empty[0]));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: both Q tiles and the K ring
    // (K and V travel through separate rings on separate warps: a K stage is free as soon as its logits are done.
    // Two query tiles per CTA share each K/V block — with one tile the 32 KiB of K/V per block were all the TMA path
    // could deliver to an SM in the time of the block's twelve MMAs.)
    if (elect_one_sync()) {
      uint32_t stage = 0, phase = 0, it = 0;
      for (int w = blockIdx.x; w < p.n_items; w += gridDim.x, ++it) {
        const int bh = w / p.q_pairs, qp = w - bh * p.q_pairs;
        bar_wait(q_empty, (it & 1u) ^ 1u);
        bar_expect(q_full, kTiles * kTile);
        const int qrow = SMALL ? w * kTQ : bh * p.sq + qp * 2 * kTQ;
        tma_2d(&map_q, q_full, sm + kOffQ, 0, qrow);
        if (kTiles == 2) tma_2d(&map_q, q_full, sm + kOffQ + kTile, 0, qrow + kTQ);
        const int krow = SMALL ? w * 2 * p.sk : bh * p.sk;
        for (int j = 0; j < p.n_kv; ++j) {
          bar_wait(k_empty + stage * 8, phase ^ 1u);
          bar_expect(k_full + stage * 8, kTile);
          tma_2d(&map_k, k_full + stage * 8, sm + kOffK + stage * kTile, 0, krow + j * kTK);
          if (++stage == kKvStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ TMA producer: the V ring
    if (elect_one_sync()) {
      uint32_t stage = 0, phase = 0;
      for (int w = blockIdx.x; w < p.n_items; w += gridDim.x) {
        const int krow = SMALL ? w * 2 * p.sk : (w / p.q_pairs) * p.sk;
        for (int j = 0; j < p.n_kv; ++j) {
          bar_wait(v_empty + stage * 8, phase ^ 1u);
          bar_expect(v_full + stage * 8, kTile);
          tma_2d(&map_v, v_full + stage * 8, sm + kOffV + stage * kTile, 0, krow + j * kTK);
          if (++stage == kKvStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer 1: logits S = Q K^T
    // Logit tiles are numbered n = 2*block + tile over the whole life of the CTA and rotate through three TMEM buffers,
    // so this thread runs up to three tiles ahead of the softmax groups.  The two products have their own issuing
    // threads: one thread issuing both spent ~190 instructions (~760 cycles) per tile, more than the tile's twelve MMAs.
    if (elect_one_sync()) {
      const uint32_t q_lo = umma_desc_lo(sm + kOffQ);
      const uint32_t kr_lo = umma_desc_lo(sm + kOffK);
      const uint32_t idesc = p.idesc_s;
      uint32_t ks = 0, kph = 0, k_lo = kr_lo;       // K ring position
      uint32_t sbuf = 0, sph = 0, d_s = tmem + kColS;   // S ring position
      uint32_t it = 0;
      for (int w = blockIdx.x; w < p.n_items; w += gridDim.x, ++it) {
        bar_wait(q_full, it & 1u);
        for (int j = 0; j < p.n_kv; ++j) {
          bar_wait(k_full + ks * 8, kph);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (t >= kTiles) break;
            bar_wait(s_empty + sbuf * 8, sph ^ 1u);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < kHD / 16; ++k)
              umma_f16_ss(d_s, umma_desc_from_lo(q_lo + t * (kTile >> 4) + 2 * k), umma_desc_from_lo(k_lo + 2 * k), idesc,
                          k != 0 ? 1u : 0u);
            commit_to(s_full + sbuf * 8);
            d_s += kTK;
            if (++sbuf == kSBufs) {
              sbuf = 0;
              sph ^= 1u;
              d_s = tmem + kColS;
            }
          }
          commit_to(k_empty + ks * 8);
          k_lo += kTile >> 4;
          if (++ks == kKvStages) {
            ks = 0;
            kph ^= 1u;
            k_lo = kr_lo;
          }
        }
        commit_to(q_empty);     // every product reading Q has been issued
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ MMA issuer 2: O += P V
    if (elect_one_sync()) {
      const uint32_t vr_lo = umma_desc_lo(sm + kOffV);
      const uint32_t p_lo = umma_desc_lo(sm + kOffP);
      const uint32_t idesc = p.idesc_o;
      uint32_t vs = 0, vph = 0, v_lo = vr_lo;       // V ring position
      uint32_t pph = 0;                             // phase of the P buffers (both tiles advance together)
      uint32_t it = 0;
      for (int w = blockIdx.x; w < p.n_items; w += gridDim.x, ++it) {
        for (int j = 0; j < p.n_kv; ++j) {
          bar_wait(v_full + vs * 8, vph);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (t >= kTiles) break;
            bar_wait(p_full + t * 8, pph);
            if (j == 0) bar_wait(o_empty + t * 8, (it & 1u) ^ 1u);     // previous item's O of this tile has been read out
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < kTK / 16; ++k)
              umma_f16_ss(tmem + kColO + t * kHD,
                          umma_desc_from_lo(p_lo + t * (2 * kTile >> 4) + (k >> 2) * (kTile >> 4) + 2 * (k & 3)),
                          umma_desc_from_lo(v_lo + k * (16 * 128 >> 4)), idesc, (j | k) != 0 ? 1u : 0u);
            commit_to(p_empty + t * 8);
            if (j + 1 == p.n_kv) commit_to(o_full + t * 8);
          }
          pph ^= 1u;
          commit_to(v_empty + vs * 8);
          v_lo += kTile >> 4;
          if (++vs == kKvStages) {
            vs = 0;
            vph ^= 1u;
            v_lo = vr_lo;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax + output: group t <-> query tile t
    const int quad = warp & 3, t = (warp - 4) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    uint32_t blk = 0, it = 0;        // key blocks done by this group over the life of the CTA
    const int ldy = p.heads * p.out_d;
    const int mine = row >> 6;       // small mode: which of the tile's two (batch, head) pairs this row belongs to
    for (int w = blockIdx.x; w < p.n_items && t < kTiles; w += gridDim.x, ++it) {
      const int bh = w / p.q_pairs, qp = w - bh * p.q_pairs;
      float l = 0.f;
      for (int j = 0; j < p.n_kv; ++j, ++blk) {
        const uint32_t n = kTiles * blk + t;           // logit tile number -> S ring slot and phase
        const uint32_t sbuf = n % kSBufs, sph = (n / kSBufs) & 1u;
        bar_wait(s_full + sbuf * 8, sph);
        tc_fence_after();
        if (p.dbg & 8) {          // handshakes only: what the TMA / MMA side sustains on its own
          tc_fence_before();
          __syncwarp();
          if (lane == 0) bar_arrive(s_empty + sbuf * 8);
          bar_wait(p_empty + t * 8, (blk & 1u) ^ 1u);
          if (lane == 0) bar_arrive(p_full + t * 8);
          continue;
        }
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          // small mode: keys [j*128 + sub*64, +64) of the item's 2 sk belong to pair 0 (< sk) or pair 1 (warp-uniform test)
          const bool other = SMALL && ((j * kTK + sub * 64 >= p.sk) ? 1 : 0) != mine;
          float v[64];
          if (!other) {
            tmem_ld32(lane_addr + kColS + sbuf * kTK + sub * 64, v);
            tmem_ld32(lane_addr + kColS + sbuf * kTK + sub * 64 + 32, v + 32);
            tmem_ld_wait();
          }
          if (sub == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) bar_arrive(s_empty + sbuf * 8);
          }
          if (other) {            // logits against the other pair's keys: P = 0, nothing added to the row sum
            if (sub == 0) bar_wait(p_empty + t * 8, (blk & 1u) ^ 1u);
            uint8_t* prow0 = smem + kOffP + (t * 2 + sub) * kTile + row * 128;
#pragma unroll
            for (int u = 0; u < 8; ++u) *reinterpret_cast<uint4*>(prow0 + (u << 4)) = make_uint4(0u, 0u, 0u, 0u);
            continue;
          }
          uint32_t pk[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            pk[i] = SCALE ? pk2(v[2 * i] * p.c1, v[2 * i + 1] * p.c1) : pk2(v[2 * i], v[2 * i + 1]);
            if (!(p.dbg & 2)) pk[i] = (POLY > 0 && i % (POLY > 0 ? POLY : 1) == POLY - 1) ? ex2_poly_pk2(pk[i]) : ex2_pk2(pk[i]);
          }
          // row sum: packed adds over 8 pairs (<= 8 * 2981 per lane, exact enough in 16 bits), then fp32
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (p.dbg & 4) break;
            uint32_t a = add_pk2(add_pk2(pk[8 * g], pk[8 * g + 1]), add_pk2(pk[8 * g + 2], pk[8 * g + 3]));
            uint32_t b = add_pk2(add_pk2(pk[8 * g + 4], pk[8 * g + 5]), add_pk2(pk[8 * g + 6], pk[8 * g + 7]));
            const float2 f = unpack_op2(add_pk2(a, b));
            l += f.x + f.y;
          }
          // P sub-tile `sub` of tile t: row of 128 bytes, 16-byte units XOR-swizzled by the row (SWIZZLE_128B)
          if (sub == 0) bar_wait(p_empty + t * 8, (blk & 1u) ^ 1u);
          uint8_t* prow = smem + kOffP + (t * 2 + sub) * kTile + row * 128;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            *reinterpret_cast<uint4*>(prow + ((u ^ (row & 7)) << 4)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) bar_arrive(p_full + t * 8);
      }
      // ---- output: O / (row sum + zero-key mass)
      const float inv = 1.0f / (l + p.zero_keys);
      bar_wait(o_full + t * 8, it & 1u);
      tc_fence_after();
      float o[64];
      tmem_ld32(lane_addr + kColO + t * kHD, o);
      tmem_ld32(lane_addr + kColO + t * kHD + 32, o + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) bar_arrive(o_empty + t * 8);
      const int bh_row = SMALL ? 2 * w + mine : bh;
      if (bh_row >= p.n_bh) continue;                       // odd number of (batch, head) pairs: the last tile is half empty
      const int b = bh_row / p.heads, h = bh_row - b * p.heads;
      const int srow = SMALL ? (row & 63) : (qp * 2 + t) * kTQ + row;
      uint4* dst = reinterpret_cast<uint4*>(p.y + (static_cast<size_t>(b) * p.sq + srow) * ldy + h * p.out_d);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (8 * u < p.out_d)
          dst[u] = make_uint4(pk2(o[8 * u] * inv, o[8 * u + 1] * inv), pk2(o[8 * u + 2] * inv, o[8 * u + 3] * inv),
                              pk2(o[8 * u + 4] * inv, o[8 * u + 5] * inv), pk2(o[8 * u + 6] * inv, o[8 * u + 7] * inv));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, kAtCols);
  }
}

}  // namespace

bool attn_tc_supported(const vb_attn_desc* d) {
  static const bool off = getenv("VB_ATTN_TC") != nullptr && atoi(getenv("VB_ATTN_TC")) == 0;     // A/B testing
  const int ld = d->ld > 0 ? d->ld : d->head_dim;
  if (off || ld != kHD || (d->head_dim != kHD && d->head_dim != 32)) return false;     // rows of 64 elements (D = 32: zero-padded)
  if (d->sq % (2 * kTQ) == 0 && d->sk % kTK == 0) return true;
  // The 8x8 level (sq == 64): two (batch, head) pairs per query tile, block-diagonal logits.  Correct and tested, but the
  // items are too short for this pipeline (11.9 us against 7.6 us for the mma.sync kernel at B=64, profiles/r02_attention.txt):
  // opt-in with VB_ATTN_SMALL=1.
  const char* e = getenv("VB_ATTN_SMALL");
  return e != nullptr && atoi(e) != 0 && d->sq == 64 && d->sk % 64 == 0 && d->head_dim == kHD;
}

int attn_tc_launch(const vb_attn_desc* d, cudaStream_t s) {
  CUtensorMap mq, mk, mv;
  const uint64_t rows_q = static_cast<uint64_t>(d->B) * d->heads * d->sq;
  const uint64_t rows_k = static_cast<uint64_t>(d->B) * d->heads * d->sk;
  const uint64_t strides[1] = {kHD * 2};
  const uint32_t box[2] = {kHD, kTQ};
  const uint64_t dq[2] = {kHD, rows_q}, dk[2] = {kHD, rows_k};
  int rc = encode_tmap_16(&mq, d->q, 2, dq, strides, box);
  if (rc != VB_OK) return rc;
  rc = encode_tmap_16(&mk, d->k, 2, dk, strides, box);
  if (rc != VB_OK) return rc;
  rc = encode_tmap_16(&mv, d->v, 2, dk, strides, box);
  if (rc != VB_OK) return rc;
  AttnTcParams p;
  memset(&p, 0, sizeof(p));
  p.small = d->sq == 64 ? 1 : 0;
  p.tiles = p.small ? 1 : 2;
  p.n_bh = d->B * d->heads;
  p.out_d = d->head_dim;
  p.q_pairs = p.small ? 1 : d->sq / (2 * kTQ);
  p.n_items = p.small ? (p.n_bh + 1) / 2 : p.n_bh * p.q_pairs;
  p.n_kv = p.small ? 2 * d->sk / kTK : d->sk / kTK;
  p.heads = d->heads;
  p.sq = d->sq;
  p.sk = d->sk;
  p.c1 = 1.4426950408889634f / sqrtf(static_cast<float>(d->head_dim));
  p.zero_keys = static_cast<float>(d->zero_keys);
  p.y = static_cast<op_t*>(d->y);
  p.idesc_s = umma_idesc_op(kTQ, kTK);
  p.idesc_o = umma_idesc_op(kTQ, kHD) | (1u << 16);        // B (= V) is MN-major: [key][d] as it lies in memory
  static const int dbg = getenv("VB_ATTN_DBG") ? atoi(getenv("VB_ATTN_DBG")) : 0;
  // share of the 2^x evaluations moved from the MUFU pipe to the FMA pipe: every POLY-th packed pair (0: none)
  static const int poly_env = getenv("VB_ATTN_POLY") ? atoi(getenv("VB_ATTN_POLY")) : 0;
  p.dbg = dbg;
  const bool scale = !(d->q_prescaled || (dbg & 1));
  typedef void (*Fn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnTcParams);
  Fn fn;
  int slot;
  if (p.small) { fn = scale ? attn_tc_kernel<true, 0, true> : attn_tc_kernel<false, 0, true>; slot = scale ? 6 : 7; }
  else if (poly_env == 2) { fn = scale ? attn_tc_kernel<true, 2, false> : attn_tc_kernel<false, 2, false>; slot = scale ? 0 : 1; }
  else if (poly_env == 3) { fn = scale ? attn_tc_kernel<true, 3, false> : attn_tc_kernel<false, 3, false>; slot = scale ? 2 : 3; }
  else { fn = scale ? attn_tc_kernel<true, 0, false> : attn_tc_kernel<false, 0, false>; slot = scale ? 4 : 5; }
  static unsigned long long attr_done[8] = {0, 0, 0, 0, 0, 0, 0, 0};        // the opt-in shared-memory attribute is per device
  const unsigned long long dev_bit = 1ull << current_device();
  if (!(attr_done[slot] & dev_bit)) {
    VB_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    attr_done[slot] |= dev_bit;
  }
  const int grid = std::min(p.n_items, num_sms());
  VB_CHECK_CUDA(launch_pdl(fn, dim3(grid), dim3(kAtThreads), kAtSmem, s, mq, mk, mv, p));
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

}  // namespace vb
