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
// CTA = 12 warps, persistent over tiles (static round-robin schedule):
//   warp 0   TMA producer (one lane)            warp 1   MMA issuer (one lane)
//   warp 2   TMEM allocator                     warps 4-11 epilogue: two warps per TMEM lane quadrant,
//                                                          each owning one 32-column half of every 64-column chunk
// Pipelines: smem full/empty ring (TMA <-> MMA), a double-buffered TMEM accumulator (MMA <->
// epilogue) so the epilogue of tile i overlaps the main loop of tile i+1, and inside the epilogue
// a residual-tile TMA ring plus an output staging ring drained by TMA stores.
//
// Epilogue structure (round-1 ncu, profiles/r01_conv_epilogue_*.txt): with K as small as 576 the main
// loop of a tile is ~1.2-1.7 k cycles, so the epilogue is co-critical.  v1 (per-thread global I/O) sat at
// L1TEX 69 %; v2 (TMA-staged I/O, 4 warps, up to three passes over TMEM, run-time switches) was issue/latency
// bound with one warp per scheduler at ~20 instructions per element.  v3 (this file): 8 warps, ONE pass over
// TMEM — the clamped result is kept packed (16-bit) in registers for the deferred pixel-norm outputs — and
// the epilogue is compiled per (residual mode, modulation, output kinds) variant so the hot loop is
// straight-line code.
#include <algorithm>
#include <cstdlib>
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
constexpr int kMaxBSlots = 12;
// Epilogue warps: PARTS per TMEM lane quadrant, each owning 64/PARTS columns of every 64-column chunk.  PARTS = 2 is the
// default; PARTS = 4 (16 epilogue warps, four per scheduler, 16 columns per thread) is a plan-time tuning choice for the
// epilogue-bound layers: the per-tile epilogue is a latency-bound instruction stream (IPC ~0.3 per scheduler with two
// warps), so more, lighter warps hide it better.
constexpr int kFirstEpiWarp = 4;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;                       // columns between the two accumulator buffers
constexpr int kStaticSmem = 5120;                     // barriers + statistic exchange (static __shared__), rounded up
constexpr int kSmemMax = 227 * 1024 - 1024 - kStaticSmem;   // dynamic smem we allow ourselves (1 KiB alignment slack)
constexpr int kMaxResSlots = 4;                       // residual TMA ring depth (16 KiB sub-tiles)
constexpr int kEpiBarrier = 1;                        // named barrier of all epilogue warps
constexpr int kPairBarrier = 3;                       // +quadrant: the warps sharing a TMEM lane quadrant (1, 2: kEpiBarrier + group)

// VB_DBG & 16: longest CTA lifetime (SM cycles) of the launches since the last vb_debug_conv_cycles() call.
__device__ unsigned long long g_conv_cycles;
// VB_DBG & 32: clock64 time stamps of CTA 0's phases (see vb_debug_conv_cycles).
__device__ long long g_conv_ts[8];
#define VB_TS(i) do { if (p.dbg & 32) { if (blockIdx.x == 0) g_conv_ts[i] = clock64(); } } while (0)
// -DVB_EPI_PROF (measurement builds only): CTA 0's epilogue leader and one other epilogue thread accumulate the cycles they
// spend in each phase of the per-chunk loop and print them when the kernel ends.
#ifdef VB_EPI_PROF
#define VB_EP_DECL long long ep_t = clock64(), ep_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const bool ep_on = blockIdx.x == 0 && (lane == 0) && (warp == kFirstEpiWarp || warp == kFirstEpiWarp + 5)
#define VB_EP(i) do { if (ep_on) { const long long n_ = clock64(); ep_acc[i] += n_ - ep_t; ep_t = n_; } } while (0)
#define VB_EP_PRINT do { if (ep_on) printf("epi warp %d leader %d: res_wait %lld tmem_ld %lld math %lld put %lld fence+drain %lld barrier %lld tma_issue %lld other %lld\n", warp, (int)leader, ep_acc[0], ep_acc[1], ep_acc[2], ep_acc[3], ep_acc[4], ep_acc[5], ep_acc[6], ep_acc[7]); } while (0)
#else
#define VB_EP_DECL
#define VB_EP(i)
#define VB_EP_PRINT
#endif

struct ConvKernelParams {
  int B, H, W;
  int bw, bh, bn;
  int tiles_x, tiles_y, tx_shift, ty_shift;
  int n_tiles, total_tiles;
  // rowroll != 0: input-stationary 3x3 for 64 -> 64 channels on rows of >= 128 pixels.  The CTA walks down a strip of
  // strip_rows output rows; every INPUT row is loaded once (one haloed box) and multiplied, with the three dx windows,
  // against [W(dy=+1) | W(dy=0) | W(dy=-1)] (N = 192, resident): its three 64-column result blocks belong to the output
  // rows above, at and below it, which occupy CONSECUTIVE 64-column slots of a ring of eight TMEM slots — so the dy sum
  // happens in the accumulator addressing, each activation row is read from shared memory 3 times instead of 9, and
  // the MMA runs at its N = 192 rate instead of the shared-memory-bound N = 64 one.  Slots are cleared by the epilogue
  // when it has read them; every MMA accumulates.
  int rowroll, strip_rows, strip_shift, chunk_mask, chunk_shift;
  uint32_t idesc128, idesc64;
  int eg;               // ping-pong epilogue (kernel variants with PARTS == 1)
  int tune_tap;         // plan-time tuning: 0 auto, 1 no shared haloed boxes, 2 haloed boxes wherever they fit
  int want_res1;        // plan-time tuning (tune bit 7): 1x1 layers keep their N tile's weights resident (see plan_mainloop)
  int pair, total_q;    // CTA-pair mode (cluster of 2, cta_group::2 MMA); work items per CTA / per pair
  int taps, kc_a, kc_b;
  int block_n;
  int num_stages, stage_bytes, b_bytes;
  // tap_mode != 0: one haloed activation box per (kc, group) serves three taps through shifted UMMA descriptor
  // windows (the SWIZZLE_128B phase is a function of the absolute smem address, so a descriptor may start at any
  // 128-byte row: tools/exp/umma_window_test.cu).  1: rows of >=128 pixels, box {bw+2} px, taps dx=-1,0,1, window
  // step 1 row.  2: multi-row tiles, box {bh+2} x {bw}, taps dy=-1,0,1, window step bw rows.
  // Two rings: a_slots activation boxes (a_slot_bytes each) and b_slots weight tap tiles (b_bytes each);
  // b_resident: the whole weight matrix of this layer stays in shared memory for the life of the CTA.
  int tap_mode, a_slots, a_slot_bytes, a_tx_bytes, b_slots, b_off, b_resident, win_rows;
  uint32_t idesc;
  int epi_mode, flags;
  const float* mod;
  int mod_stride;
  int res_mode;         // VB_RES_*
  int res_slots;        // residual ring depth (power of two)
  int nslots;           // staged output slots in use
  int out_kind[3];
  float out_scale[3];
  int gslots;           // sub-tiles per staging region (a region = the outputs of one chunk in one pass)
  int stg_regions;      // staging regions in the ring (3: the store of the previous region may still be reading; 2: it may not)
  int res_off, stg_off; // byte offsets of the residual ring / staging ring inside dynamic smem
  float* out_f32;
  int ld_f32;
  float* out_rnorm;          // per-pixel 1/(eps + rms) side channel (producer)
  const float* res_rnorm;    // ... and its consumer (VB_RES_SCALED)
  float res_a, res_b, clamp;
  float inv_sqrt_c;     // 1/sqrt(cout) for the pixel norms
  int head_dim, parts, seg_div, heads;
  op_t* part0;
  op_t* part1;
  op_t* part2;
  int part_seq[3];
  int part_off[3];
  float norm_scale;   // 1/sqrt(head_dim)
  int part_ld;          // QKVNORM: elements per destination row (head_dim, or 64 for zero-padded D = 32 rows)
  float part_scale[3];  // QKVNORM: multiplier of the normalised rows of part j (q: log2(e)/sqrt(D) of the softmax)
  int dbg;            // ablation switches for micro-benchmarks (VB_DBG): 1 = epilogue does no work, 2 = no TMA stores,
                      // 16 = record CTA lifetimes
  int qkv_stg;        // QKVNORM, D = 64: the normalised (token, head, q|k|v) rows leave through a shared-memory staging ring and
                      // TMA stores (map_out.m[part]: {64, rows} 2-D maps of the three destinations) instead of per-thread stores
  int qkv_rows;       // ... rows per store box = min(128, H*W)
  int wgt_nowait;     // set per launch: the weight producer need not wait for the previous kernel (see the PDL note in the kernel)
  // K-split (tune bit 8): the two CTAs of a cluster work on the SAME tile, each over half of the input-channel blocks of
  // every tap.  Rank 1 writes its fp32 partial accumulator to ks_ws ([tile][block_n/4][128 rows][4], L2-resident) and
  // publishes a per-warp tile counter in rank 0's shared memory; rank 0 adds the partial to its own accumulator values
  // (always own + partner: a fixed order) and runs the normal epilogue.  For layers with fewer tiles than SMs and long K
  // loops (the 8x8 level).
  int ksplit;
  float* ks_ws;
};

struct TileCoord {
  int x0, y0, n0, col0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvKernelParams& p, int tile) {
  // tiles_x and tiles_y are powers of two (H, W and the tile extents are); only n_tiles needs a real division
  const int mt = p.n_tiles == 1 ? tile : tile / p.n_tiles;
  const int nt = tile - mt * p.n_tiles;
  const int tx = mt & (p.tiles_x - 1);
  const int t2 = mt >> p.tx_shift;
  const int ty = t2 & (p.tiles_y - 1);
  const int tn = t2 >> p.ty_shift;
  TileCoord t;
  t.x0 = tx * p.bw;
  t.y0 = ty * p.bh;
  t.n0 = tn * p.bn;
  t.col0 = nt * p.block_n;
  return t;
}

// Two floats -> one packed 16-bit pair, saturating to the finite range (no inf in the stream).
__device__ __forceinline__ uint32_t pack_sat2(float lo, float hi) {
#ifdef VB_OP_BF16
  return pack_op2(lo, hi);
#else
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
#endif
}

// Plain round-to-nearest pack for values already clamped to the finite range.
__device__ __forceinline__ uint32_t pack_op2_nosat(float lo, float hi) {
#ifdef VB_OP_BF16
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
#else
  __half2 h = __floats2half2_rn(lo, hi);
#endif
  return *reinterpret_cast<uint32_t*>(&h);
}

// mp_silu of a packed 16-bit pair, computed in packed arithmetic: x*c*(1 + tanh(x/2)), c = 0.5/0.596 — one MUFU op and
// three packed multiplies per TWO elements (the fp32 form costs five instructions and one MUFU op per element; the
// epilogue of the narrow SR layers is issue-bound).  Same error class as tanh.approx.f32 followed by 16-bit rounding.
__device__ __forceinline__ uint32_t mp_silu_pk(uint32_t x2, float scale) {
  // y = (x c)(1 + tanh(x/2)) with x = scale * x2: both scalings ride on constants — h = x2 (scale/2), xc = x2 (scale c),
  // y = fma(xc, tanh(h), xc): three packed operations and the tanh per pair
#ifdef VB_OP_BF16
  const __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&x2);
  __nv_bfloat162 h = __hmul2(x, __float2bfloat162_rn(0.5f * scale));
  const __nv_bfloat162 xc = __hmul2(x, __float2bfloat162_rn(scale * (0.5f / 0.596f)));
  uint32_t hu = *reinterpret_cast<uint32_t*>(&h), tu;
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(tu) : "r"(hu));
  __nv_bfloat162 y = __hfma2(xc, *reinterpret_cast<__nv_bfloat162*>(&tu), xc);
#else
  const __half2 x = *reinterpret_cast<__half2*>(&x2);
  __half2 h = __hmul2(x, __float2half2_rn(0.5f * scale));
  const __half2 xc = __hmul2(x, __float2half2_rn(scale * (0.5f / 0.596f)));
  uint32_t hu = *reinterpret_cast<uint32_t*>(&h), tu;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(tu) : "r"(hu));
  __half2 y = __hfma2(xc, *reinterpret_cast<__half2*>(&tu), xc);
#endif
  return *reinterpret_cast<uint32_t*>(&y);
}

// Clamp of a packed fp16 pair to [-c, c] (c exactly representable in fp16).
__device__ __forceinline__ uint32_t clamp_pk(uint32_t x2, float c) {
#ifdef VB_OP_BF16
  __nv_bfloat162 cc = __float2bfloat162_rn(c);
  __nv_bfloat162 y = __hmin2(__hmax2(*reinterpret_cast<__nv_bfloat162*>(&x2), __hneg2(cc)), cc);
#else
  __half2 cc = __float2half2_rn(c);
  __half2 y = __hmin2(__hmax2(*reinterpret_cast<__half2*>(&x2), __hneg2(cc)), cc);
#endif
  return *reinterpret_cast<uint32_t*>(&y);
}

// Byte offset of 16-byte unit `u` (0..7) of a 128-byte row in a SWIZZLE_128B sub-tile.
__device__ __forceinline__ int swz(int u, int row) { return (u ^ (row & 7)) << 4; }

// Sum of squares of this row's 32 residual columns [half*32, half*32+32) (VB_RES_PIXNORM statistic).
template <int U>
__device__ __forceinline__ float res_sumsq32(const uint8_t* rrow, int row, int part) {
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const uint4 q = *reinterpret_cast<const uint4*>(rrow + swz(part * U + j, row));
    const float2 a = unpack_op2(q.x), b = unpack_op2(q.y), c = unpack_op2(q.z), d = unpack_op2(q.w);
    ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
  }
  return ss;
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
  float s0 = 0.f, s1 = 0.f;             // (even / odd partial sums: the same order as the staged form below)
#pragma unroll
  for (int j = 0; j < D; j += 2) fma2(s0, s1, v[j], v[j + 1], v[j], v[j + 1]);
  const float ss = s0 + s1;
  const int gg = gcol / D;
  const int part = gg % p.parts;
  const float inv = (part == 0 ? p.part_scale[0] : (part == 1 ? p.part_scale[1] : p.part_scale[2])) /
                    (1e-4f + sqrtf(ss) * p.norm_scale);
  const int head = gg / p.parts;
  const int b = n / p.seg_div;
  const int seg = n - b * p.seg_div;
  op_t* base = part == 0 ? p.part0 : (part == 1 ? p.part1 : p.part2);
  const int seq = part == 0 ? p.part_seq[0] : (part == 1 ? p.part_seq[1] : p.part_seq[2]);
  const int off = part == 0 ? p.part_off[0] : (part == 1 ? p.part_off[1] : p.part_off[2]);
  const size_t tok = (static_cast<size_t>(b) * p.heads + head) * seq + off + seg * (p.H * p.W) + s;
  uint4* o = reinterpret_cast<uint4*>(base + tok * p.part_ld);
#pragma unroll
  for (int j = 0; j < D / 8; ++j)
    o[j] = make_uint4(pack_op2(v[8 * j] * inv, v[8 * j + 1] * inv), pack_op2(v[8 * j + 2] * inv, v[8 * j + 3] * inv),
                      pack_op2(v[8 * j + 4] * inv, v[8 * j + 5] * inv), pack_op2(v[8 * j + 6] * inv, v[8 * j + 7] * inv));
}

// The same for D = 64 through shared memory: the thread writes its normalised 128-byte row into a SWIZZLE_128B staging
// sub-tile (conflict-free 16-byte stores); one thread per column group then hands the 128 rows to a TMA store.  (Written by
// the thread itself, a row costs eight store instructions that each touch 32 different 128-byte lines, 16 bytes apiece:
// ~3 k LSU cycles per 128x192 tile, as long as the tile's main loop.)
__device__ __forceinline__ void epi_group_qkv_stage(const ConvKernelParams& p, uint32_t taddr, int gcol, uint8_t* srow, int row) {
  float v[64];
  tmem_ld32(taddr, v);
  tmem_ld32(taddr + 32, v + 32);
  tmem_ld_wait();
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 64; j += 2) fma2(s0, s1, v[j], v[j + 1], v[j], v[j + 1]);
  const float ss = s0 + s1;
  const int part = (gcol >> 6) % p.parts;
  const float inv = (part == 0 ? p.part_scale[0] : (part == 1 ? p.part_scale[1] : p.part_scale[2])) /
                    (1e-4f + sqrtf(ss) * p.norm_scale);
#pragma unroll
  for (int j = 0; j < 64; j += 2) mul2(v[j], v[j + 1], inv, inv);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<uint4*>(srow + swz(j, row)) =
        make_uint4(pack_op2(v[8 * j], v[8 * j + 1]), pack_op2(v[8 * j + 2], v[8 * j + 3]),
                   pack_op2(v[8 * j + 4], v[8 * j + 5]), pack_op2(v[8 * j + 6], v[8 * j + 7]));
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

// Linear tile index (M tile major over N tiles) of work item q for CTA `rank`.  Single CTA: q itself.  CTA pair: the
// pair takes M tiles 2m and 2m+1 of one N tile; an odd tail decodes to an image index >= B, where TMA zero-fills the
// loads and drops the stores.
__device__ __forceinline__ int tile_of(const ConvKernelParams& p, int q, uint32_t rank) {
  if (!p.pair) return q;
  if (p.n_tiles == 1) return 2 * q + static_cast<int>(rank);
  const int mq = q / p.n_tiles;
  return (2 * mq + static_cast<int>(rank)) * p.n_tiles + (q - mq * p.n_tiles);
}

// rowroll: output row i (0..strip_rows-1; -2,-1 address the two leading garbage blocks) of strip q.  Strips are
// numbered x tile fastest, then chunk of rows, then image.
__device__ __forceinline__ TileCoord strip_tile(const ConvKernelParams& p, int q, int i) {
  TileCoord t;
  t.x0 = (q & (p.tiles_x - 1)) * p.bw;
  const int t2 = q >> p.tx_shift;
  t.y0 = ((t2 & p.chunk_mask) << p.strip_shift) + i;
  t.n0 = t2 >> p.chunk_shift;
  t.col0 = 0;
  return t;
}

// ------------------------------------------------------------------------------------------------ single-thread roles
// The TMA producers and the MMA issuer are ONE thread each; a lone warp issues a dependent instruction every ~5 cycles,
// so their instruction streams are the pipeline's clock: the first version (one producer thread, barrier addresses
// re-derived from the shared-window base and %cluster_ctarank at every use, run-time branches on the layer's mode)
// spent ~65 instructions (~450 cycles) per 64-channel K block — slower than the four MMAs of the block for every tile
// narrower than N=256 (ncu source view, profiles/r01_conv_roles.txt).  Hence: barrier and shared-memory addresses are
// plain 32-bit values computed once, the operand loads are split over two warps (activations: warp 0, weights: warp 3),
// the pair / single-CTA instruction forms are template arguments, and ring positions advance by adds.
struct RoleCtx {
  uint32_t smem;        // shared::cta address of the (1024-byte aligned) dynamic shared memory
  uint32_t full, empty, bfull, bempty, tmem_full, tmem_empty;   // local barrier arrays (shared::cta addresses)
  uint32_t full_dst, bfull_dst;                                  // where TMA completes its bytes (pair: leader CTA's)
  uint32_t tmem_base;
  uint32_t rank;
  int q0, qstride;
  int k_lo, k_hi;                 // this CTA's range of 64-channel K blocks within every tap (K-split: half of them)
  int ka_lo, ka_hi, kb_lo, kb_hi; // ... as ranges of the first / second channel-concatenated source
};

// Pins a computed address in a register: without it the compiler re-derives every barrier address from the shared
// window base and SR_CgaCtaId (5 instructions) at each use inside the single-thread loops.
__device__ __forceinline__ uint32_t pinned(uint32_t v) {
  uint32_t r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity) {
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
    if (++spins > VB_SPIN_LIMIT) __trap();       // a stuck pipeline traps instead of hanging the GPU
  }
}
__device__ __forceinline__ void mbar_expect_tx_addr(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
template <bool PAIR>
__device__ __forceinline__ void tma_act(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
  if (PAIR)
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
        "%6}], [%2];" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_act(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0),
               "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
template <bool PAIR>
__device__ __forceinline__ void tma_wgt(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1) {
  if (PAIR)
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  else
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
template <bool PAIR>
__device__ __forceinline__ void umma_commit_addr(uint32_t bar) {
  if (PAIR)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-split hand-over (see ConvKernelParams::ksplit): the partner's epilogue warp w publishes "partials of my first n tiles are in
// the workspace" by a release store into THIS CTA's shared memory; warp w here polls its own counter.  A counter instead of an
// mbarrier: the producer may run several tiles ahead and a phase bit would wrap.
__device__ __forceinline__ void ks_publish(uint32_t counter_cluster_addr, uint32_t n) {
  asm volatile("st.release.cluster.shared::cluster.u32 [%0], %1;" ::"r"(counter_cluster_addr), "r"(n) : "memory");
}
__device__ __forceinline__ void ks_wait(const uint32_t* counter, uint32_t n) {
  const uint32_t a = smem_u32(counter);
  uint32_t spins = 0;
  while (true) {
    uint32_t v;
    asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    if (v >= n) return;
    __nanosleep(64);
    if (++spins > (1u << 24)) __trap();          // a stuck partner traps instead of hanging the GPU
  }
}

// Activation producer (warp 0).  Simple mode: one {64 ch, tile} box per (tap, K chunk) stage; it also arms the stage's
// barrier for the weight tile the other producer sends.  Tap mode: one haloed box per (tap group, K chunk).
template <bool PAIR>
__device__ __forceinline__ void producer_act(const ConvKernelParams& p, const RoleCtx& c, const CUtensorMap* map_a,
                                             const CUtensorMap* map_a2) {
  const bool lead = c.rank == 0;
  const uint32_t mul = PAIR ? 2u : 1u;
  uint32_t slot = 0, phase = 0;
  if (p.rowroll) {
    if (p.dbg & 4) return;                                  // ablation: no operand loads (MMA on stale shared memory)
    const uint32_t tx = static_cast<uint32_t>(p.a_tx_bytes);
    const uint32_t nslots = p.a_slots, slot_bytes = p.a_slot_bytes;
    uint32_t dst = c.smem;
    for (int q = c.q0; q < p.total_q; q += c.qstride) {
      const TileCoord t = strip_tile(p, q, 0);
      constexpr int kAhead = 8;                             // rows pulled into L2 ahead of the shared-memory ring: every
      for (int j = 0; j < kAhead; ++j)                      // row is read from HBM exactly once, so without this the ring
        tma_prefetch_act(map_a, 0, t.x0 - 1, t.y0 - 1 + j, t.n0);      // (5-6 rows) has to cover the full DRAM latency
      for (int j = 0; j < p.strip_rows + 2; ++j) {          // input rows y0-1 .. y0+strip_rows, each loaded ONCE
        if (j + kAhead < p.strip_rows + 2) tma_prefetch_act(map_a, 0, t.x0 - 1, t.y0 - 1 + j + kAhead, t.n0);
        mbar_wait_addr(c.empty + slot * 8, phase ^ 1u);
        mbar_expect_tx_addr(c.full + slot * 8, tx);
        tma_act<false>(map_a, c.full + slot * 8, dst, 0, t.x0 - 1, t.y0 - 1 + j, t.n0);
        dst += slot_bytes;
        if (++slot == nslots) {
          slot = 0;
          phase ^= 1u;
          dst = c.smem;
        }
      }
    }
  } else if (p.tap_mode != 0) {
    const uint32_t tx = static_cast<uint32_t>(p.a_tx_bytes) * mul;
    const uint32_t nslots = p.a_slots, slot_bytes = p.a_slot_bytes;
    uint32_t dst = c.smem;
    for (int q = c.q0; q < p.total_q; q += c.qstride) {
      const TileCoord t = decode_tile(p, tile_of(p, q, c.rank));
      for (int g = 0; g < 3; ++g) {
        const int cx = p.tap_mode == 1 ? t.x0 - 1 : t.x0 + g - 1;
        const int cy = p.tap_mode == 1 ? t.y0 + g - 1 : t.y0 - 1;
        for (int half = 0; half < 2; ++half) {
          const CUtensorMap* m = half == 0 ? map_a : map_a2;
          const int n = half == 0 ? c.ka_hi : c.kb_hi;
          for (int kc = half == 0 ? c.ka_lo : c.kb_lo; kc < n; ++kc) {
            mbar_wait_addr(c.empty + slot * 8, phase ^ 1u);
            if (lead) mbar_expect_tx_addr(c.full + slot * 8, tx);
            tma_act<PAIR>(m, c.full_dst + slot * 8, dst, kc * kBlockK, cx, cy, t.n0);
            dst += slot_bytes;
            if (++slot == nslots) {
              slot = 0;
              phase ^= 1u;
              dst = c.smem;
            }
          }
        }
      }
    }
  } else {
    // (resident 1x1 weights: the stage carries the activation box only)
    const uint32_t tx = static_cast<uint32_t>(kAStageBytes + (p.b_resident ? 0 : p.b_bytes)) * mul;
    const uint32_t nslots = p.num_stages, slot_bytes = p.stage_bytes;
    uint32_t dst = c.smem;
    for (int q = c.q0; q < p.total_q; q += c.qstride) {
      const TileCoord t = decode_tile(p, tile_of(p, q, c.rank));
      int dy = p.taps == 9 ? -1 : 0, dx = dy;
      for (int tap = 0; tap < p.taps; ++tap) {
        for (int half = 0; half < 2; ++half) {
          const CUtensorMap* m = half == 0 ? map_a : map_a2;
          const int n = half == 0 ? c.ka_hi : c.kb_hi;
          for (int kc = half == 0 ? c.ka_lo : c.kb_lo; kc < n; ++kc) {
            mbar_wait_addr(c.empty + slot * 8, phase ^ 1u);
            if (lead) mbar_expect_tx_addr(c.full + slot * 8, tx);
            tma_act<PAIR>(m, c.full_dst + slot * 8, dst, kc * kBlockK, t.x0 + dx, t.y0 + dy, t.n0);
            dst += slot_bytes;
            if (++slot == nslots) {
              slot = 0;
              phase ^= 1u;
              dst = c.smem;
            }
          }
        }
        if (++dx == 2) {
          dx = -1;
          ++dy;
        }
      }
    }
  }
}

// Weight producer (warp 3).  Simple mode: the {64, rows} tile of every stage (the stage barrier was armed by the
// activation producer; complete_tx may run ahead of expect_tx within a phase).  Tap mode: its own ring of tap tiles, or
// the whole layer once when it is resident.
template <bool PAIR>
__device__ __forceinline__ void producer_wgt(const ConvKernelParams& p, const RoleCtx& c, const CUtensorMap* map_w) {
  const bool lead = c.rank == 0;
  const uint32_t mul = PAIR ? 2u : 1u;
  const int wrow = PAIR ? static_cast<int>(c.rank) * (p.block_n >> 1) : 0;     // this CTA's rows of the weight tile
  const int kct = p.kc_a + p.kc_b;
  const uint32_t b_bytes = p.b_bytes;
  if (p.rowroll) {
    // B tile of window dx: rows [0,64) = W(dy=+1), [64,128) = W(dy=0), [128,192) = W(dy=-1), each a {64 k, 64 rows} box
    mbar_expect_tx_addr(c.bfull, 9u * b_bytes);
    for (int dx = 0; dx < 3; ++dx)
      for (int kb = 0; kb < 3; ++kb)
        tma_wgt<false>(map_w, c.bfull, c.smem + p.b_off + (dx * 3 + kb) * b_bytes, ((2 - kb) * 3 + dx) * kBlockK, 0);
    return;
  }
  if (p.tap_mode != 0) {
    if (p.b_resident) {
      const int nb = 9 * kct;
      if (lead) mbar_expect_tx_addr(c.bfull, static_cast<uint32_t>(nb) * b_bytes * mul);
      uint32_t dst = c.smem + p.b_off;
      for (int i = 0; i < nb; ++i, dst += b_bytes) tma_wgt<PAIR>(map_w, c.bfull_dst, dst, i * kBlockK, wrow);
      return;
    }
    const uint32_t nslots = p.b_slots, base = c.smem + p.b_off;
    const uint32_t tx = b_bytes * mul;
    // K column of (g, i, kc): (tap * kct + kc) * 64 with tap = 3g+i (mode 1) or 3i+g (mode 2)
    const int g_step = (p.tap_mode == 1 ? 3 * kct : kct) * kBlockK;
    const int i_step = (p.tap_mode == 1 ? kct : 3 * kct) * kBlockK;
    uint32_t slot = 0, phase = 0, dst = base;
    for (int q = c.q0; q < p.total_q; q += c.qstride) {
      const int col = decode_tile(p, tile_of(p, q, c.rank)).col0 + wrow;
      for (int g = 0; g < 3; ++g) {
        for (int kc = c.k_lo; kc < c.k_hi; ++kc) {
          int kcol = g * g_step + kc * kBlockK;
#pragma unroll
          for (int i = 0; i < 3; ++i, kcol += i_step) {
            mbar_wait_addr(c.bempty + slot * 8, phase ^ 1u);
            if (lead) mbar_expect_tx_addr(c.bfull + slot * 8, tx);
            tma_wgt<PAIR>(map_w, c.bfull_dst + slot * 8, dst, kcol, col);
            dst += b_bytes;
            if (++slot == nslots) {
              slot = 0;
              phase ^= 1u;
              dst = base;
            }
          }
        }
      }
    }
  } else if (p.b_resident) {
    // 1x1 layer, resident weights: the grid is a multiple of the number of N tiles, so every work item of this CTA has
    // the SAME N tile; its K x block_n weight slab is loaded once and the ring carries activations only
    if (c.q0 >= p.total_q) return;
    const int col = decode_tile(p, tile_of(p, c.q0, c.rank)).col0 + wrow;
    if (lead) mbar_expect_tx_addr(c.bfull, static_cast<uint32_t>(kct) * b_bytes * mul);
    uint32_t dst = c.smem + p.b_off;
    for (int i = 0; i < kct; ++i, dst += b_bytes) tma_wgt<PAIR>(map_w, c.bfull_dst, dst, i * kBlockK, col);
  } else {
    const uint32_t nslots = p.num_stages, slot_bytes = p.stage_bytes, base = c.smem + kAStageBytes;
    const int k_end = p.taps * kct;
    uint32_t slot = 0, phase = 0, dst = base;
    for (int q = c.q0; q < p.total_q; q += c.qstride) {
      const int col = decode_tile(p, tile_of(p, q, c.rank)).col0 + wrow;
      for (int kb0 = 0; kb0 < k_end; kb0 += kct) {             // per tap: this CTA's K blocks (all of them unless K-split)
        for (int kb = kb0 + c.k_lo; kb < kb0 + c.k_hi; ++kb) {
          mbar_wait_addr(c.empty + slot * 8, phase ^ 1u);
          tma_wgt<PAIR>(map_w, c.full_dst + slot * 8, dst, kb * kBlockK, col);
          dst += slot_bytes;
          if (++slot == nslots) {
            slot = 0;
            phase ^= 1u;
            dst = base;
          }
        }
      }
    }
  }
}

// The MMA issuer (warp 1; pair: the leader CTA's).  Descriptors advance by adds on the low word.
template <bool PAIR>
__device__ __forceinline__ void mma_role(const ConvKernelParams& p, const RoleCtx& c) {
  const uint32_t idesc = p.idesc;
  auto mma = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t acc) {
    if (PAIR) umma_f16_ss_pair(d, umma_desc_from_lo(a_lo), umma_desc_from_lo(b_lo), idesc, acc);
    else umma_f16_ss(d, umma_desc_from_lo(a_lo), umma_desc_from_lo(b_lo), idesc, acc);
  };
  const int kct = p.kc_a + p.kc_b;
  const int kcl = c.k_hi - c.k_lo;               // K blocks per tap this CTA multiplies (kct unless K-split; never with resident weights)
  const uint32_t b_tile_lo = static_cast<uint32_t>(p.b_bytes) >> 4;
  int it = 0;
  if (p.rowroll) {
    const uint32_t a_base = umma_desc_lo(c.smem);
    const uint32_t a_slot_lo = static_cast<uint32_t>(p.a_slot_bytes) >> 4;
    const uint32_t b_base = umma_desc_lo(c.smem + p.b_off);
    const uint32_t b_dx_lo = 3u * b_tile_lo;                  // one window's 192-row weight tile
    const uint32_t i128 = p.idesc128, i64 = p.idesc64;
    const bool no_mma = (p.dbg & 8) != 0;                      // ablation: loads and handshakes only
    auto mma_n = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t id) {
      if (!no_mma) umma_f16_ss(d, umma_desc_from_lo(a_lo), umma_desc_from_lo(b_lo), id, 1u);
    };
    const uint32_t a_slots = p.a_slots;
    uint32_t as = 0, aph = 0, a_lo = a_base;
    uint32_t g = 0;                                           // input rows processed so far = first of its three slots
    mbar_wait_addr(c.bfull, 0);
    mbar_wait_addr(c.tmem_empty, 0);                          // slots 0 and 1 of the very first row: the initial clear
    mbar_wait_addr(c.tmem_empty + 8, 0);
    tc_fence_after();
    for (int q = c.q0; q < p.total_q; q += c.qstride) {
      for (int j = 0; j < p.strip_rows + 2; ++j, ++g) {
        const uint32_t nb = g + 2;                            // the block this row starts: its slot must have been cleared
        mbar_wait_addr(c.tmem_empty + (nb & 7u) * 8, (nb >> 3) & 1u);
        if (!(p.dbg & 4)) mbar_wait_addr(c.full + as * 8, aph);
        if (g == 0) VB_TS(2);
        tc_fence_after();
        const uint32_t s0 = g & 7u;
        const uint32_t d0 = c.tmem_base + s0 * 64u;
        // (the slot case is decided once per row: decided per MMA it cost ~35 instructions per MMA in the issuing thread)
        if (s0 <= 5u) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) mma_n(d0, a_lo + i * 8 + 2 * k, b_base + i * b_dx_lo + 2 * k, idesc);
          }
        } else if (s0 == 6u) {                                // slots 6,7 then wrap to 0
#pragma unroll
          for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              mma_n(d0, a_lo + i * 8 + 2 * k, b_base + i * b_dx_lo + 2 * k, i128);
              mma_n(c.tmem_base, a_lo + i * 8 + 2 * k, b_base + i * b_dx_lo + 2 * k + 2u * b_tile_lo, i64);
            }
          }
        } else {                                              // slot 7 then wrap to 0,1
#pragma unroll
          for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              mma_n(d0, a_lo + i * 8 + 2 * k, b_base + i * b_dx_lo + 2 * k, i64);
              mma_n(c.tmem_base, a_lo + i * 8 + 2 * k, b_base + i * b_dx_lo + 2 * k + b_tile_lo, i128);
            }
          }
        }
        umma_commit_addr<false>(c.empty + as * 8);
        umma_commit_addr<false>(c.tmem_full + s0 * 8);        // block g has received its last contribution
        a_lo += a_slot_lo;
        if (++as == a_slots) {
          as = 0;
          aph ^= 1u;
          a_lo = a_base;
        }
      }
    }
    VB_TS(3);
  } else if (p.tap_mode != 0) {
    const uint32_t a_base = umma_desc_lo(c.smem);
    const uint32_t a_slot_lo = static_cast<uint32_t>(p.a_slot_bytes) >> 4;
    const uint32_t win_lo = static_cast<uint32_t>(p.win_rows) * 8u;          // rows of 128 B
    const uint32_t b_base = umma_desc_lo(c.smem + p.b_off);
    // resident weights: tile index of (g, i, kc) is (tap * kct + kc) with tap = 3g+i (mode 1) or 3i+g (mode 2)
    const uint32_t b_g_lo = static_cast<uint32_t>(p.tap_mode == 1 ? 3 * kct : kct) * b_tile_lo;
    const uint32_t b_i_lo = static_cast<uint32_t>(p.tap_mode == 1 ? kct : 3 * kct) * b_tile_lo;
    const uint32_t a_slots = p.a_slots, b_slots = p.b_slots;
    uint32_t as = 0, bs = 0, aph = 0, bph = 0, a_lo = a_base, b_ring = b_base;
    const bool resident = p.b_resident != 0;
    if (resident) {
      mbar_wait_addr(c.bfull, 0);
      tc_fence_after();
    }
    for (int q = c.q0; q < p.total_q; q += c.qstride, ++it) {
      const uint32_t buf = it & 1;
      mbar_wait_addr(c.tmem_empty + buf * 8, (static_cast<uint32_t>(it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = c.tmem_base + buf * kAccStride;
      uint32_t accumulate = 0;
      for (int g = 0; g < 3; ++g) {
        uint32_t b_gk = b_base + static_cast<uint32_t>(g) * b_g_lo;
        for (int kc = 0; kc < kcl; ++kc, b_gk += b_tile_lo) {
          mbar_wait_addr(c.full + as * 8, aph);
          if (it == 0 && g == 0 && kc == 0) VB_TS(2);
          tc_fence_after();
          if (resident) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                mma(d_tmem, a_lo + i * win_lo + 2 * k, b_gk + i * b_i_lo + 2 * k, accumulate);
                accumulate = 1;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              mbar_wait_addr(c.bfull + bs * 8, bph);
              tc_fence_after();
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                mma(d_tmem, a_lo + i * win_lo + 2 * k, b_ring + 2 * k, accumulate);
                accumulate = 1;
              }
              umma_commit_addr<PAIR>(c.bempty + bs * 8);
              b_ring += b_tile_lo;
              if (++bs == b_slots) {
                bs = 0;
                bph ^= 1u;
                b_ring = b_base;
              }
            }
          }
          umma_commit_addr<PAIR>(c.empty + as * 8);
          a_lo += a_slot_lo;
          if (++as == a_slots) {
            as = 0;
            aph ^= 1u;
            a_lo = a_base;
          }
        }
      }
      umma_commit_addr<PAIR>(c.tmem_full + buf * 8);
      VB_TS(3);
    }
  } else {
    const uint32_t s_base = umma_desc_lo(c.smem);
    const uint32_t stage_lo = static_cast<uint32_t>(p.stage_bytes) >> 4;
    const uint32_t nstages = p.num_stages;
    uint32_t stage = 0, phase = 0, a_lo = s_base;
    const int k_blocks = p.taps * kcl;
    const bool resident = p.b_resident != 0;
    const uint32_t b_res = umma_desc_lo(c.smem + p.b_off);
    if (resident && c.q0 < p.total_q) {
      mbar_wait_addr(c.bfull, 0);
      tc_fence_after();
    }
    for (int q = c.q0; q < p.total_q; q += c.qstride, ++it) {
      const uint32_t buf = it & 1;
      mbar_wait_addr(c.tmem_empty + buf * 8, (static_cast<uint32_t>(it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = c.tmem_base + buf * kAccStride;
      uint32_t accumulate = 0;
      uint32_t b_next = b_res;
      for (int kb = 0; kb < k_blocks; ++kb, b_next += b_tile_lo) {
        mbar_wait_addr(c.full + stage * 8, phase);
        if (it == 0 && kb == 0) VB_TS(2);
        tc_fence_after();
        const uint32_t b_lo = resident ? b_next : a_lo + (kAStageBytes >> 4);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          mma(d_tmem, a_lo + 2 * k, b_lo + 2 * k, accumulate);
          accumulate = 1;
        }
        umma_commit_addr<PAIR>(c.empty + stage * 8);   // smem slot reusable once these MMAs retire
        a_lo += stage_lo;
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1u;
          a_lo = s_base;
        }
      }
      umma_commit_addr<PAIR>(c.tmem_full + buf * 8);       // accumulator complete -> epilogue
      VB_TS(3);
    }
  }
}

struct OutMaps {
  CUtensorMap m[3];
};

__device__ __forceinline__ bool kind_direct(int k) { return k == VB_OUT_RAW || k == VB_OUT_SILU; }
__device__ __forceinline__ bool kind_norm(int k) { return k == VB_OUT_NORM || k == VB_OUT_NORM_SILU; }

// Template arguments >= 0 fix the epilogue variant at compile time; -1 reads it from the parameters (generic
// fallback for combinations the plans never emit).  STAGED = 0 is the QKVNORM / narrow fp32 epilogue.
// KS: the variant carries the K-split epilogue paths (only the variants the 8x8-level 3x3 layers use are instantiated with it;
// the others stay instruction-for-instruction what they were — several sit at the 168-register limit).
template <int STAGED, int RES_T, int MOD_T, int K0_T, int K1_T, int K2_T, int PARTS, bool KS = false>
__global__ void __launch_bounds__(PARTS == 1 ? 384 : 128 + 128 * PARTS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                 const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_res,
                 const __grid_constant__ OutMaps map_out, const __grid_constant__ ConvKernelParams p) {
  // PARTS == 1 selects the ping-pong epilogue (EG): two groups of four warps take ALTERNATE tiles (group g <-> accumulator
  // buffer g), a thread owns a whole 64-column row (two 32-column halves in sequence), each group has its own residual
  // ring, staging ring, barrier and TMA-issuing thread — so the latency chain of one tile's epilogue (accumulator
  // wake-up, tcgen05.ld, proxy fence, staging barrier, TMA issue: ~2 k cycles even for the plainest epilogue) overlaps
  // the other group's.  For tiles of one or two chunks (block_n <= 128).
  constexpr bool EG = PARTS == 1;
  constexpr int NP = EG ? 2 : PARTS;      // column parts of a chunk (EG: processed in sequence by one thread)
  constexpr int HH = EG ? 2 : 1;
  constexpr int MAXC = EG ? 2 : 4;        // chunks per tile
  constexpr int kEpiWarps = EG ? 8 : 4 * PARTS;
  constexpr int kTileWarps = EG ? 4 : kEpiWarps;      // warps working on one tile
  constexpr int kEpiThreads = kTileWarps * 32;
  // specialised variant without pixel-norm outputs (nothing is kept across the chunks of a tile)
  constexpr bool kNoNormT = K0_T >= 0 && K1_T >= 0 && K2_T >= 0 && K0_T < VB_OUT_NORM && K1_T < VB_OUT_NORM && K2_T < VB_OUT_NORM;
  constexpr int CW = 64 / NP;             // accumulator columns per thread per part
  constexpr int U = CW / 8;               // 16-byte units per part per chunk row
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t bfull_bar[kMaxBSlots];
  __shared__ __align__(8) uint64_t bempty_bar[kMaxBSlots];
  __shared__ __align__(8) uint64_t tmem_full[8];      // two accumulator buffers, or (rowroll) eight 64-column slots
  __shared__ __align__(8) uint64_t tmem_empty[8];
  __shared__ __align__(8) uint64_t res_full[kMaxResSlots];
  __shared__ __align__(8) uint64_t res_empty[kMaxResSlots];
  __shared__ float xchg[2][NP][kBlockM];     // [residual | result statistic][column part][row]
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t ks_count[8];           // K-split: tiles whose partial the partner's epilogue warp i has published

  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long t_start = (p.dbg & 16) ? clock64() : 0;
  if (threadIdx.x == 0) VB_TS(0);
  // CTA pair (p.pair): the two CTAs of a cluster work on two adjacent 128-pixel tiles of the same output-channel tile.
  // Each loads its own activation operand and HALF of the weight tile; the leader (rank 0) issues one M=256
  // cta_group::2 MMA for both, so every CTA reads (128 + block_n/2) operand rows per K step from its shared memory
  // instead of (128 + block_n) — the SS-mode MMA is shared-memory-bandwidth bound below N=256 (tools/exp/umma_rate.cu).
  const uint32_t rank = p.pair ? cluster_ctarank() : 0u;
  // K-split: the two CTAs of a cluster take the SAME work items and half of every tap's K blocks each (otherwise independent:
  // own operands, own MMA issuer, own accumulator)
  const uint32_t krank = p.ksplit ? cluster_ctarank() : 0u;
  const bool shared_items = p.pair || p.ksplit;
  const int q0 = shared_items ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int qstride = shared_items ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    if (p.kc_b > 0) tma_prefetch_desc(&map_a2);
    if (p.res_mode != VB_RES_NONE) tma_prefetch_desc(&map_res);
    for (int s = 0; s < p.nslots; ++s) tma_prefetch_desc(&map_out.m[s]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kMaxBSlots; ++s) {
      mbar_init(&bfull_bar[s], 1);
      mbar_init(&bempty_bar[s], 1);
    }
    for (int b = 0; b < 8; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], p.pair ? 2 * kTileWarps : kTileWarps);   // pair: the epilogue warps of both CTAs
    }
    for (int b = 0; b < kMaxResSlots; ++b) {
      mbar_init(&res_full[b], 1);
      mbar_init(&res_empty[b], kTileWarps);
    }
    for (int b = 0; b < 8; ++b) ks_count[b] = 0;
    fence_mbar_init();
  }
  if (warp == 2) {
    if (p.pair) {
      tmem_alloc_pair(&tmem_slot, kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(&tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (shared_items) cluster_sync_all(); else __syncthreads();      // (K-split: the partner's counters are zeroed before it publishes)
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // Everything above overlapped the previous kernel's tail; dependent global memory is touched from here on.  The weight
  // producer (warp 3) does NOT wait: prepared weights are constants of the plan (written at plan build, long before any
  // replay), so its first tiles — or the whole resident slab — arrive while the previous kernel is still draining.
  // (Only when the launch says the previous kernel of the stream is another op of the same plan, p.wgt_nowait: a one-shot
  // vb_conv, or the first op of a replayed range, may directly follow the kernel that wrote its weights.)
  if (warp != 3 || !p.wgt_nowait) pdl_grid_sync(); else pdl_launch_dependents();
  if (threadIdx.x == 0) VB_TS(1);

  if (warp < 2 || warp == 3) {
    // ------------------------------------------------------------------ producers (warps 0, 3) and MMA issuer (warp 1)
    if (elect_one_sync()) {
      RoleCtx c;
      c.smem = pinned(smem_u32(smem));
      c.full = pinned(smem_u32(&full_bar[0]));
      c.empty = pinned(smem_u32(&empty_bar[0]));
      c.bfull = pinned(smem_u32(&bfull_bar[0]));
      c.bempty = pinned(smem_u32(&bempty_bar[0]));
      c.tmem_full = pinned(smem_u32(&tmem_full[0]));
      c.tmem_empty = pinned(smem_u32(&tmem_empty[0]));
      // pair: both CTAs complete their bytes on the LEADER's full barriers, which the leader arms for both halves
      c.full_dst = pinned(p.pair ? mapa_u32(c.full, 0) : c.full);
      c.bfull_dst = pinned(p.pair ? mapa_u32(c.bfull, 0) : c.bfull);
      c.tmem_base = tmem_base;
      c.rank = rank;
      c.q0 = q0;
      c.qstride = qstride;
      {
        const int kct = p.kc_a + p.kc_b;
        c.k_lo = p.ksplit ? static_cast<int>(krank) * (kct >> 1) : 0;
        c.k_hi = p.ksplit ? c.k_lo + (kct >> 1) : kct;
        c.ka_lo = min(c.k_lo, p.kc_a);
        c.ka_hi = min(c.k_hi, p.kc_a);
        c.kb_lo = max(c.k_lo - p.kc_a, 0);
        c.kb_hi = max(c.k_hi - p.kc_a, 0);
      }
      if (warp == 0) {
        if (p.pair) producer_act<true>(p, c, &map_a, &map_a2); else producer_act<false>(p, c, &map_a, &map_a2);
      } else if (warp == 3) {
        if (p.pair) producer_wgt<true>(p, c, &map_w); else producer_wgt<false>(p, c, &map_w);
      } else if (rank == 0) {
        if (p.pair) mma_role<true>(p, c); else mma_role<false>(p, c);
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ------------------------------------------------------------------ epilogue
    const int quad = warp & 3;                                      // TMEM lane quadrant this warp may read
    const int wq = (warp - kFirstEpiWarp) >> 2;
    const int group = EG ? wq : 0;                                  // EG: which of the two alternating tile streams
    const int half = EG ? 0 : wq;                                   // which CW columns of every 64-column chunk (part)
    const int row = quad * 32 + lane;
    // one thread per group issues its TMA traffic (staged QKVNORM epilogue: one per column part)
    const bool leader = elect_one_sync() && warp == kFirstEpiWarp + 4 * ((!STAGED && p.qkv_stg) ? wq : group);
    VB_EP_DECL;
    const int rx = row % p.bw;
    const int r2 = row / p.bw;
    const int ry = r2 % p.bh;
    const int rn = r2 / p.bh;
    int it = 0;
    // accumulator buffer drained: tell the MMA issuer (pair: the leader CTA's barrier counts the warps of both CTAs)
    const uint32_t tmem_empty_leader = p.pair ? mapa_u32(smem_u32(&tmem_empty[0]), 0) : 0u;
    auto acc_release = [&](int buf) {
      if (p.pair) mbar_arrive_cluster_addr(tmem_empty_leader + buf * 8);
      else mbar_arrive(&tmem_empty[buf]);
    };

    if (!STAGED) {
      uint32_t qkv_g = 0;               // staged QKVNORM: groups this warp has processed (staging slot = parity)
      for (int q = q0; q < p.total_q; q += qstride, ++it) {
        const int buf = it & 1;
        const uint32_t bphase = static_cast<uint32_t>(it >> 1) & 1u;
        const TileCoord t = decode_tile(p, tile_of(p, q, rank));
        const int n = t.n0 + rn;
        const bool valid = n < p.B;
        const int s_img = (t.y0 + ry) * p.W + t.x0 + rx;
        const size_t pix = static_cast<size_t>(n) * p.H * p.W + s_img;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(buf * kAccStride);
        mbar_wait(&tmem_full[buf], bphase);
        if (leader && it == 0) VB_TS(4);
        tc_fence_after();
        if (p.epi_mode == VB_EPI_QKVNORM && p.qkv_stg) {
          // Each column part (the four warps `half` of the four lane quadrants) works through its own groups with its own
          // two staging slots, named barrier and store-issuing thread; groups rotate over the parts as below.
          const int first = (half + it) % PARTS;
          uint8_t* stg = smem + p.stg_off + half * 2 * kChunkBytes;
          for (int c = first * 64; c < p.block_n; c += 64 * PARTS) {
            uint8_t* slot = stg + (qkv_g & 1u) * kChunkBytes;
            epi_group_qkv_stage(p, taddr + c, t.col0 + c, slot + row * 128, row);
            fence_proxy_async();
            // the store issued two groups ago read THIS group's next slot; it must be done before anybody passes the barrier
            if (leader) bulk_wait_read<0>();
            named_bar_sync(kEpiBarrier + half, 128);
            if (leader) {
              const int gg = (t.col0 + c) >> 6;
              const int part = gg % p.parts, head = gg / p.parts;
              const int seq = part == 0 ? p.part_seq[0] : (part == 1 ? p.part_seq[1] : p.part_seq[2]);
              const int off = part == 0 ? p.part_off[0] : (part == 1 ? p.part_off[1] : p.part_off[2]);
              const int hw = p.H * p.W;
              for (int j = 0; j < p.bn; ++j) {                  // one box per image of the tile (bn = 2 at 8x8)
                const int nj = t.n0 + j;
                if (nj >= p.B) break;
                const int b = nj / p.seg_div, seg = nj - b * p.seg_div;
                const int tok = (b * p.heads + head) * seq + off + seg * hw + (p.bn == 1 ? t.y0 * p.W + t.x0 : 0);
                tma_store_2d(&map_out.m[part], slot + j * p.qkv_rows * 128, 0, tok);
              }
              bulk_commit();
            }
            ++qkv_g;
          }
        } else if (p.epi_mode == VB_EPI_QKVNORM) {
          // the column groups of a tile are dealt to the PARTS warps of a lane quadrant starting with a different warp on
          // every tile: with an odd number of groups (block_n = 192, D = 64) the extra one alternates instead of always
          // landing on the same warp (2 : 1 became 3 : 3 over two tiles)
          const int first = (half + it) % PARTS;
          if (p.head_dim == 64) {
            for (int c = first * 64; c < p.block_n; c += 64 * PARTS) epi_group_qkv<64>(p, taddr + c, t.col0 + c, n, s_img, valid);
          } else {
            for (int c = first * 32; c < p.block_n; c += 32 * PARTS) epi_group_qkv<32>(p, taddr + c, t.col0 + c, n, s_img, valid);
          }
        } else {
          for (int c = half * 16; c < p.block_n; c += 16 * PARTS) epi_f32_16(p, taddr + c, t.col0 + c, pix, valid);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) acc_release(buf);
      }
      if (p.qkv_stg && leader) bulk_wait_read<0>();      // staging smem must outlive the last TMA store's read
    } else if (KS && !EG && p.ksplit && krank == 1) {
      // K-split, second half of K: this CTA's accumulator is a PARTIAL sum.  Each epilogue thread writes its row's columns as
      // fp32 into the tile's workspace block — laid out [column / 4][row][4] so that a warp's 16-byte stores (32 consecutive
      // rows) are one contiguous 512-byte run — and the warp then publishes its tile count to the partner warp of rank 0,
      // which owns exactly the same rows and columns.
      const uint32_t counter = mapa_u32(smem_u32(&ks_count[warp - kFirstEpiWarp]), 0);
      const int chunks = p.block_n >> 6;
      for (it = 0;; ++it) {
        const int q = q0 + it * qstride;
        if (q >= p.total_q) break;
        const int buf = it & 1;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(buf * kAccStride + half * CW);
        float4* ws = reinterpret_cast<float4*>(p.ks_ws + static_cast<size_t>(q) * (kBlockM * p.block_n)) + half * (CW / 4) * kBlockM + row;
        mbar_wait(&tmem_full[buf], static_cast<uint32_t>(it >> 1) & 1u);
        tc_fence_after();
        for (int c = 0; c < chunks; ++c) {
          float v[CW];
          if (CW == 32) tmem_ld32(taddr + c * 64, v); else tmem_ld16(taddr + c * 64, v);
          tmem_ld_wait();
          if (c == chunks - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) acc_release(buf);
          }
          float4* o = ws + c * 16 * kBlockM;
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) __stcg(o + j * kBlockM, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) ks_publish(counter, static_cast<uint32_t>(it + 1));
      }
    } else {
      const int res_mode = RES_T >= 0 ? RES_T : p.res_mode;
      const bool rowroll = !EG && p.rowroll != 0;          // (the ping-pong variants never run the row-rolling layout: compile-time false)
      const bool modsilu = MOD_T >= 0 ? MOD_T != 0 : (p.flags & VB_F_MODSILU) != 0;
      const int k0 = K0_T >= 0 ? K0_T : p.out_kind[0];
      const int k1 = K1_T >= 0 ? K1_T : p.out_kind[1];
      const int k2 = K2_T >= 0 ? K2_T : p.out_kind[2];
      const bool has_res = res_mode != VB_RES_NONE;
      const bool needs_norm = kind_norm(k0) || kind_norm(k1) || kind_norm(k2);
      const bool any_direct = kind_direct(k0) || kind_direct(k1) || kind_direct(k2);
      const int chunks = p.block_n >> 6;                            // 64-column sub-tiles
      uint8_t* res_ring = smem + p.res_off + group * p.res_slots * kChunkBytes;
      uint8_t* stg_ring = smem + p.stg_off + group * p.stg_regions * p.gslots * kChunkBytes;
      uint64_t* res_full_g = res_full + group * 2;            // EG: two slots per group
      uint64_t* res_empty_g = res_empty + group * 2;
      const uint32_t rmask = static_cast<uint32_t>(p.res_slots - 1);
      const uint32_t rshift = p.res_slots == 4 ? 2u : 1u;
      const int items = chunks * (res_mode == VB_RES_PIXNORM ? 2 : 1);   // residual (pass, chunk) items per tile
      uint32_t res_q = 0;          // residual items consumed so far by this thread (ring position / phase)
      uint32_t res_issued = 0;     // residual items whose TMA load has been issued (leader only)
      uint32_t greg = 0;           // staging region the next group of sub-tiles goes to (ring position)
      const float clampv = p.clamp;

      // The leader keeps the residual ring full ACROSS tile boundaries: the next tile's residual is in flight while
      // this tile is being finished.
      // (the leader's instruction stream is serial and sits on every chunk's critical path: the tile of the next residual
      //  item is decoded once per tile and the item index advances by adds — no divisions per item)
      uint32_t rt_item = 0;                      // item within the tile whose residual is issued next
      uint32_t rt_tile = EG ? group : 0;         // CTA-local index of that tile
      bool rt_have = false;
      TileCoord rt_t = {0, 0, 0, 0};
      auto res_topup = [&]() {
        if (has_res && leader) {
          while (res_issued < res_q + static_cast<uint32_t>(p.res_slots)) {
            if (!rt_have) {
              const int wl = rowroll ? static_cast<int>(rt_tile >> p.strip_shift) : static_cast<int>(rt_tile);
              const int ql = q0 + wl * qstride;
              if (ql >= p.total_q) break;
              rt_t = rowroll ? strip_tile(p, ql, static_cast<int>(rt_tile) & (p.strip_rows - 1))
                               : decode_tile(p, tile_of(p, ql, rank));
              rt_have = true;
            }
            const uint32_t slot = res_issued & rmask;
            mbar_wait(&res_empty_g[slot], ((res_issued >> rshift) & 1u) ^ 1u);
            mbar_expect_tx(&res_full_g[slot], kChunkBytes);
            const int rc = static_cast<int>(rt_item >= static_cast<uint32_t>(chunks) ? rt_item - chunks : rt_item);
            tma_load_4d(&map_res, &res_full_g[slot], res_ring + slot * kChunkBytes, rt_t.col0 + rc * 64, rt_t.x0, rt_t.y0,
                        rt_t.n0);
            ++res_issued;
            if (++rt_item == static_cast<uint32_t>(items)) {
              rt_item = 0;
              rt_tile += EG ? 2 : 1;
              rt_have = false;
            }
          }
        }
      };
      // Makes the next residual item readable; returns this row's 128-byte slot of it.
      auto res_acquire = [&]() -> const uint8_t* {
        res_topup();
        mbar_wait(&res_full_g[res_q & rmask], (res_q >> rshift) & 1u);
        return res_ring + (res_q & rmask) * kChunkBytes + row * 128;
      };
      auto res_release = [&]() {
        __syncwarp();
        if (lane == 0) mbar_arrive(&res_empty_g[res_q & rmask]);
        ++res_q;
      };
      // Region of the staging ring the next group of sub-tiles goes to.  With three regions the leader only has to know
      // that the store issued TWO groups ago has finished reading its region (the one the next group will overwrite)
      // before it lets everybody past the barrier — the most recent store stays in flight.  (With a two-region ring and
      // a wait for ALL stores, 14 % of the kernel's stall samples sat on this barrier: profiles/r01_conv_stage_ring.txt.)
      auto stg_region = [&]() -> uint8_t* { return stg_ring + greg * p.gslots * kChunkBytes; };
      // All 256 threads wrote their part of the region: publish it to the async proxy and let the leader store it.
      auto stg_commit = [&](const TileCoord& t, int c, bool norm_pass, int only = -1) {
        VB_EP(3);
        fence_proxy_async();
        if (leader) {
          if (p.stg_regions == 3) bulk_wait_read<1>(); else bulk_wait_read<0>();
        }
        VB_EP(4);
        named_bar_sync(kEpiBarrier + group, kEpiThreads);
        VB_EP(5);
        if (leader && !(p.dbg & 2)) {
          const uint8_t* reg = stg_region();
          int di = 0;
          if ((only < 0 || only == 0) && (norm_pass ? kind_norm(k0) : kind_direct(k0))) tma_store_4d(&map_out.m[0], reg + (di++) * kChunkBytes, t.col0 + c * 64, t.x0, t.y0, t.n0);
          if ((only < 0 || only == 1) && (norm_pass ? kind_norm(k1) : kind_direct(k1))) tma_store_4d(&map_out.m[1], reg + (di++) * kChunkBytes, t.col0 + c * 64, t.x0, t.y0, t.n0);
          if ((only < 0 || only == 2) && (norm_pass ? kind_norm(k2) : kind_direct(k2))) tma_store_4d(&map_out.m[2], reg + (di++) * kChunkBytes, t.col0 + c * 64, t.x0, t.y0, t.n0);
          bulk_commit();
        }
        VB_EP(6);
        if (++greg == static_cast<uint32_t>(p.stg_regions)) greg = 0;
      };

      // rowroll: clears this warp's lanes x columns of accumulator slot `col` (every MMA of that mode accumulates)
      auto slot_clear = [&](uint32_t taddr_slot) {
        if (p.dbg & 64) return;                 // ablation: no clearing (wrong results)
        tmem_st32_fill(taddr_slot, 0u);
        tmem_st_wait();
      };
      const int nblk = rowroll ? p.strip_rows + 2 : 1;            // accumulator blocks per work item
      if (rowroll) {                                              // all eight slots start cleared and free
        for (int sl = 0; sl < 8; ++sl) {
          slot_clear(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(sl * 64 + half * CW));
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[sl]);
        }
      }
      const bool dbg_off = (p.dbg & 1) != 0, dbg_ts = (p.dbg & 32) != 0;       // ablation / time-stamp switches, read once
      for (it = EG ? group : 0;; it += EG ? 2 : 1) {                // ping-pong: group g takes tiles g, g+2, ...
        const int wi = rowroll ? it / nblk : it;                  // CTA-local work item (tile, or strip of rows)
        const int jb = rowroll ? it - wi * nblk : 0;              // block within the strip: 0,1 carry no output row
        const int q = q0 + wi * qstride;
        if (q >= p.total_q) break;
        const int buf = rowroll ? (it & 7) : (it & 1);
        const uint32_t bphase = static_cast<uint32_t>(rowroll ? it >> 3 : it >> 1) & 1u;
        const TileCoord t = rowroll ? strip_tile(p, q, jb - 2) : decode_tile(p, tile_of(p, q, rank));
        const int n = t.n0 + rn;
        const bool valid = n < p.B;
        // (pixel index: evaluated where it is used — the per-pixel side channel and the fp32 output — so that variants with
        //  neither do not carry its 64-bit arithmetic through every tile)
        auto pix_of = [&]() -> size_t { return static_cast<size_t>(n) * p.H * p.W + (t.y0 + ry) * p.W + t.x0 + rx; };
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                               static_cast<uint32_t>((rowroll ? buf * 64 : buf * kAccStride) + half * CW);
        if (rowroll && jb < 2) {                                  // partial sums of rows outside the strip: discard
          mbar_wait(&tmem_full[buf], bphase);
          tc_fence_after();
          slot_clear(taddr);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) acc_release(buf);
          continue;
        }
        if (!dbg_off) res_topup();

        if (dbg_off) {
          mbar_wait(&tmem_full[buf], bphase);
          tc_fence_after();
          if (rowroll) slot_clear(taddr);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) acc_release(buf);
          continue;
        }

        // ---- pass R: pixel-norm statistic of the residual row (VB_RES_PIXNORM)
        float res_scale = p.res_a;
        if (res_mode == VB_RES_SCALED) res_scale = p.res_a * __ldg(p.res_rnorm + (valid ? pix_of() : 0));
        if (res_mode == VB_RES_PIXNORM) {
          float ss = 0.f;
          for (int c = 0; c < chunks; ++c) {
            const uint8_t* rrow = res_acquire();
            ss += res_sumsq32<U>(rrow, row, half);
            res_release();
          }
          xchg[0][half][row] = ss;
          named_bar_sync(kPairBarrier + quad, 32 * NP);       // (never reached by the ping-pong variants)
#pragma unroll
          for (int o = 1; o < NP; ++o) ss += xchg[0][(half + o) % NP][row];
          res_scale = p.res_a / (1e-4f + sqrtf(ss) * p.inv_sqrt_c);
        }

        // modulation row of this (image, column block): fetched ahead of the accumulator wait, the next chunk's while the
        // current one is packed and stored (L2 latency off the per-tile critical path)
        float4 mreg[CW / 4];
        auto mod_fetch = [&](int c, int part) {
          const float4* m = reinterpret_cast<const float4*>(p.mod + static_cast<size_t>(valid ? n : 0) * p.mod_stride + t.col0 +
                                                            c * 64 + part * CW);
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) mreg[j] = __ldg(m + j);
        };
        if (modsilu) mod_fetch(0, half);

        VB_EP(7);
        mbar_wait(&tmem_full[buf], bphase);
#ifdef VB_EPI_PROF
        if (ep_on) ep_t = clock64();          // (waiting for the accumulator is not epilogue work)
#endif
        if (dbg_ts && leader && it == 0) VB_TS(4);
        tc_fence_after();

        // ---- pass M: accumulator -> modulation / mp_silu -> mp_sum with the residual -> clamp; RAW / SILU outputs leave
        // now, the clamped value stays packed in registers for the pixel-norm outputs.
        uint32_t keep[MAXC][HH][CW / 2];
        float ssp[HH];                              // sum of squares per column part (summed in the same order in every layout)
#pragma unroll
        for (int hh = 0; hh < HH; ++hh) ssp[hh] = 0.f;
        // One 64-column chunk of pass M.  (A lambda so that the chunk loop can be rolled or unrolled, see below.)
        auto pass_m = [&](const int c) {
          {
            VB_EP(7);
            const uint8_t* rrow = has_res ? res_acquire() : nullptr;
            VB_EP(0);
            uint32_t r16h[HH][CW / 2];
#pragma unroll
            for (int hh = 0; hh < HH; ++hh) {
              const int part = EG ? hh : half;
              const uint32_t ta = taddr + static_cast<uint32_t>(c * 64 + (EG ? hh * CW : 0));
              if (EG && modsilu && (c > 0 || hh > 0)) mod_fetch(c, part);
              float v[CW];
              if (CW == 32) tmem_ld32(ta, v); else tmem_ld16(ta, v);
              tmem_ld_wait();
              if (KS && !EG && p.ksplit) {
                // K-split: add the partner CTA's partial sum of this tile (second half of K; L2-resident workspace) — as own + partner
                // in every launch alike.  (Fetched after the accumulator read, not prefetched into registers beside it: these
                // variants sit at the 168-register limit, and the layers that use this are main-loop bound.)
                if (c == 0) {
                  if (lane == 0) ks_wait(&ks_count[warp - kFirstEpiWarp], static_cast<uint32_t>(it + 1));
                  __syncwarp();
                  __threadfence();
                }
                const float4* w4 = reinterpret_cast<const float4*>(p.ks_ws + static_cast<size_t>(q) * (kBlockM * p.block_n)) +
                                   (c * 16 + part * (CW / 4)) * kBlockM + row;
#pragma unroll
                for (int j = 0; j < CW / 4; ++j) {
                  const float4 a = __ldcg(w4 + j * kBlockM);
                  v[4 * j + 0] += a.x;
                  v[4 * j + 1] += a.y;
                  v[4 * j + 2] += a.z;
                  v[4 * j + 3] += a.w;
                }
              }
              VB_EP(1);
              if (c == chunks - 1 && hh == HH - 1) {   // accumulator fully read: the MMA warp may start the tile after next
                if (rowroll) slot_clear(taddr);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) acc_release(buf);
              }
              const int col = t.col0 + c * 64 + part * CW;
              // conv_res0: v = mp_silu(v * mod) in packed fp32 arithmetic — per column PAIR two multiplies, two tanh, one multiply and
              // one FMA (seven instructions with the pack; the packed-fp16 form needed eight with its half-register shuffle, and
              // rounded before the activation instead of after it)
              if (modsilu) {
#pragma unroll
                for (int j = 0; j < CW / 4; ++j) {
                  const float4 mm = mreg[j];
                  mul2(v[4 * j + 0], v[4 * j + 1], mm.x, mm.y);
                  mul2(v[4 * j + 2], v[4 * j + 3], mm.z, mm.w);
                  mp_silu_f32x2(v[4 * j + 0], v[4 * j + 1]);
                  mp_silu_f32x2(v[4 * j + 2], v[4 * j + 3]);
                }
                if (!EG && c + 1 < chunks) mod_fetch(c + 1, half);
              }
              if (has_res) {
                // mp_sum(res, v, t) = res * a + v * b.  Plans fold b into the prepared weights (VB_F_RESB_FOLDED: one multiply per
                // element less in every residual layer's epilogue); the descriptor form with plain weights keeps it here.
                if (p.flags & VB_F_RESB_FOLDED) {
#pragma unroll
                  for (int j = 0; j < U; ++j) {
                    const uint4 q = *reinterpret_cast<const uint4*>(rrow + swz(part * U + j, row));
                    const float2 a = unpack_op2(q.x), b = unpack_op2(q.y), cc = unpack_op2(q.z), d = unpack_op2(q.w);
                    fma2(v[8 * j + 0], v[8 * j + 1], a.x, a.y, res_scale, res_scale);
                    fma2(v[8 * j + 2], v[8 * j + 3], b.x, b.y, res_scale, res_scale);
                    fma2(v[8 * j + 4], v[8 * j + 5], cc.x, cc.y, res_scale, res_scale);
                    fma2(v[8 * j + 6], v[8 * j + 7], d.x, d.y, res_scale, res_scale);
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < U; ++j) {
                    const uint4 q = *reinterpret_cast<const uint4*>(rrow + swz(part * U + j, row));
                    const float2 a = unpack_op2(q.x), b = unpack_op2(q.y), cc = unpack_op2(q.z), d = unpack_op2(q.w);
                    v[8 * j + 0] = fmaf(a.x, res_scale, v[8 * j + 0] * p.res_b);
                    v[8 * j + 1] = fmaf(a.y, res_scale, v[8 * j + 1] * p.res_b);
                    v[8 * j + 2] = fmaf(b.x, res_scale, v[8 * j + 2] * p.res_b);
                    v[8 * j + 3] = fmaf(b.y, res_scale, v[8 * j + 3] * p.res_b);
                    v[8 * j + 4] = fmaf(cc.x, res_scale, v[8 * j + 4] * p.res_b);
                    v[8 * j + 5] = fmaf(cc.y, res_scale, v[8 * j + 5] * p.res_b);
                    v[8 * j + 6] = fmaf(d.x, res_scale, v[8 * j + 6] * p.res_b);
                    v[8 * j + 7] = fmaf(d.y, res_scale, v[8 * j + 7] * p.res_b);
                  }
                }
              }
#ifdef VB_OP_BF16
              constexpr bool kClampPacked = false;
#else
              // fp16 stream without pixel-norm outputs: the saturating pack replaces the +-65504 clamp, and the clip of
              // Block.forward (a bound that fp16 represents exactly, so clip-then-round == round-then-clip) runs on the packed
              // pairs — 16 (or 0) instructions per 32 columns instead of 64
              constexpr bool kClampPacked = kNoNormT;
#endif
              if (!kClampPacked || p.out_f32 != nullptr) {
#pragma unroll
                for (int j = 0; j < CW; ++j) v[j] = fminf(fmaxf(v[j], -clampv), clampv);
              }
              if (needs_norm) {          // two partial sums (even / odd columns) in one packed FMA per pair — in every layout alike
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int j = 0; j < CW; j += 2) fma2(s0, s1, v[j], v[j + 1], v[j], v[j + 1]);
                ssp[hh] += s0 + s1;
              }
              if (p.out_f32 != nullptr && valid) {
                float4* o = reinterpret_cast<float4*>(p.out_f32 + pix_of() * p.ld_f32 + col);
#pragma unroll
                for (int j = 0; j < CW / 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              }
              if (kClampPacked) {
#pragma unroll
                for (int j = 0; j < CW / 2; ++j) r16h[hh][j] = pack_sat2(v[2 * j], v[2 * j + 1]);
                if (p.flags & VB_F_CLIP) {
#pragma unroll
                  for (int j = 0; j < CW / 2; ++j) r16h[hh][j] = clamp_pk(r16h[hh][j], clampv);
                }
              } else {
#pragma unroll
                for (int j = 0; j < CW / 2; ++j) r16h[hh][j] = pack_op2_nosat(v[2 * j], v[2 * j + 1]);
              }
              if (needs_norm) {
#pragma unroll
                for (int j = 0; j < CW / 2; ++j) keep[c][hh][j] = r16h[hh][j];
              }
            }
            if (has_res) res_release();
            VB_EP(2);
            if (any_direct) {
              // one output after the other through a single staging slot (ping-pong epilogue: its two rings must fit), or
              // all outputs of the chunk side by side and one commit
              auto put = [&](int kind, float scale, uint8_t* srow) {
#pragma unroll
                for (int hh = 0; hh < HH; ++hh) {
                  const int part = EG ? hh : half;
#pragma unroll
                  for (int j = 0; j < U; ++j) {
                    uint4 o;
                    if (kind == VB_OUT_RAW) {
                      o = make_uint4(r16h[hh][4 * j], r16h[hh][4 * j + 1], r16h[hh][4 * j + 2], r16h[hh][4 * j + 3]);
                    } else {
                      o.x = mp_silu_pk(r16h[hh][4 * j + 0], scale);
                      o.y = mp_silu_pk(r16h[hh][4 * j + 1], scale);
                      o.z = mp_silu_pk(r16h[hh][4 * j + 2], scale);
                      o.w = mp_silu_pk(r16h[hh][4 * j + 3], scale);
                    }
                    *reinterpret_cast<uint4*>(srow + swz(part * U + j, row)) = o;
                  }
                }
              };
              if (EG) {
                if (kind_direct(k0)) { put(k0, p.out_scale[0], stg_region() + row * 128); stg_commit(t, c, false, 0); }
                if (kind_direct(k1)) { put(k1, p.out_scale[1], stg_region() + row * 128); stg_commit(t, c, false, 1); }
                if (kind_direct(k2)) { put(k2, p.out_scale[2], stg_region() + row * 128); stg_commit(t, c, false, 2); }
              } else {
                uint8_t* srow = stg_region() + row * 128;
                if (kind_direct(k0)) { put(k0, p.out_scale[0], srow); srow += kChunkBytes; }
                if (kind_direct(k1)) { put(k1, p.out_scale[1], srow); srow += kChunkBytes; }
                if (kind_direct(k2)) { put(k2, p.out_scale[2], srow); srow += kChunkBytes; }
                stg_commit(t, c, false);
              }
            }
          }
        };
        // Variants without pixel-norm outputs keep nothing across chunks, so their chunk loop stays ROLLED: the epilogue of a
        // tile is then ~0.6 k instructions instead of MAXC x that.  It matters because the short layers (16x16 / 8x8: one or two
        // tiles per CTA) execute this code once or twice per launch, straight out of L2: with four unrolled copies (35-64 KiB
        // per tile pass against a 32 KiB L1.5 / 6 KiB L0 instruction cache) half of the epilogue warps' stall samples on the
        // 1x1 attn_proj layers were instruction fetches (stall_no_inst, profiles/r02_conv_icache.txt).
        if (kNoNormT) {
#pragma unroll 1
          for (int c = 0; c < chunks; ++c) pass_m(c);
        } else {
#pragma unroll
          for (int c = 0; c < MAXC; ++c) {
            if (c < chunks) pass_m(c);
          }
        }

        // ---- pass N: pixel-norm outputs from the packed registers
        if (needs_norm) {
          float ssv;
          if (EG) {
            ssv = ssp[0] + ssp[HH - 1];
          } else {
            ssv = ssp[0];
            xchg[1][half][row] = ssv;
            named_bar_sync(kPairBarrier + quad, 32 * NP);
#pragma unroll
            for (int o = 1; o < NP; ++o) ssv += xchg[1][(half + o) % NP][row];
          }
          const float inv_v = 1.0f / (1e-4f + sqrtf(ssv) * p.inv_sqrt_c);
          if (p.out_rnorm != nullptr && half == 0 && valid) p.out_rnorm[pix_of()] = inv_v;
#pragma unroll
          for (int c = 0; c < MAXC; ++c) {
            if (c < chunks) {
              auto putn = [&](int kind, uint8_t* srow) {
#pragma unroll
                for (int hh = 0; hh < HH; ++hh) {
                  const int part = EG ? hh : half;
#pragma unroll
                  for (int j = 0; j < U; ++j) {
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      if (kind == VB_OUT_NORM) {
                        const float2 x = unpack_op2(keep[c][hh][4 * j + e]);
                        o[e] = pack_sat2(x.x * inv_v, x.y * inv_v);
                      } else {
                        o[e] = mp_silu_pk(keep[c][hh][4 * j + e], inv_v);
                      }
                    }
                    *reinterpret_cast<uint4*>(srow + swz(part * U + j, row)) = make_uint4(o[0], o[1], o[2], o[3]);
                  }
                }
              };
              if (EG) {
                if (kind_norm(k0)) { putn(k0, stg_region() + row * 128); stg_commit(t, c, true, 0); }
                if (kind_norm(k1)) { putn(k1, stg_region() + row * 128); stg_commit(t, c, true, 1); }
                if (kind_norm(k2)) { putn(k2, stg_region() + row * 128); stg_commit(t, c, true, 2); }
              } else {
                uint8_t* srow = stg_region() + row * 128;
                if (kind_norm(k0)) { putn(k0, srow); srow += kChunkBytes; }
                if (kind_norm(k1)) { putn(k1, srow); srow += kChunkBytes; }
                if (kind_norm(k2)) { putn(k2, srow); srow += kChunkBytes; }
                stg_commit(t, c, true);
              }
            }
          }
        }
      }
      if (leader) bulk_wait_read<0>();          // staging smem must outlive the last TMA store's read
      if (leader) VB_TS(5);
      VB_EP_PRINT;
    }
  }

  tc_fence_before();
  if (p.pair) cluster_sync_all(); else __syncthreads();   // pair: the peer's barriers / TMEM stay valid until both are done
  if (warp == 2) {
    tc_fence_after();
    if (p.pair) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
  if (threadIdx.x == 0) VB_TS(6);
  if ((p.dbg & 16) && threadIdx.x == 0) atomicMax(&g_conv_cycles, static_cast<unsigned long long>(clock64() - t_start));
}

}  // namespace

typedef void (*ConvKernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const OutMaps,
                             const ConvKernelParams);

struct Variant {
  int staged, res, mod, k0, k1, k2;     // -1 = any (run-time switch inside the kernel)
  int parts;                            // epilogue warps per TMEM lane quadrant (2, or 4 as a plan-time tuning choice)
  int ks;                               // carries the K-split paths
  ConvKernelFn fn;
  unsigned long long attr_done;   // bit per device ordinal: the opt-in shared-memory attribute is per device
};
#define VB_VARIANT(S, R, M, A, B, C, P) {S, R, M, A, B, C, P, 0, conv_gemm_kernel<S, R, M, A, B, C, P>, 0ull}
#define VB_VARIANT_KS(S, R, M, A, B, C) {S, R, M, A, B, C, 2, 1, conv_gemm_kernel<S, R, M, A, B, C, 2, true>, 0ull}
// (the 16-warp form, PARTS = 4, is kept compilable — add VB_VARIANT(..., 4) here — but not instantiated: measured on B200 it
//  was 1-5 % SLOWER on every epilogue-bound layer, profiles/r01_conv_epilogue_notes.txt, so the epilogue is not bound by
//  per-warp latency)
#define VB_VARIANT24(S, R, M, A, B, C) VB_VARIANT(S, R, M, A, B, C, 2), VB_VARIANT(S, R, M, A, B, C, 1)
// The epilogue combinations the plans emit (engine.py) get straight-line code; anything else runs the generic one.
static Variant g_variants[] = {
    VB_VARIANT_KS(1, 0, 1, VB_OUT_RAW, 0, 0),                              // K-split conv_res0 / conv_res1 of the 8x8 level
    VB_VARIANT_KS(1, 1, 0, VB_OUT_RAW, 0, 0),
    VB_VARIANT(0, -1, -1, -1, -1, -1, 2),                                  // QKVNORM / narrow fp32
    VB_VARIANT24(1, 0, 1, VB_OUT_RAW, 0, 0),                               // conv_res0: modulation + mp_silu
    VB_VARIANT24(1, 0, 0, VB_OUT_RAW, 0, 0),                               // conv_skip, first conv
    VB_VARIANT24(1, 0, 0, VB_OUT_RAW, VB_OUT_NORM_SILU, 0),
    VB_VARIANT24(1, 0, 0, VB_OUT_RAW, VB_OUT_SILU, 0),
    VB_VARIANT24(1, 0, 0, VB_OUT_RAW, VB_OUT_NORM_SILU, VB_OUT_SILU),
    VB_VARIANT24(1, 0, 0, VB_OUT_NORM, VB_OUT_NORM_SILU, 0),               // enc conv_skip + pixel-norm
    VB_VARIANT24(1, 1, 0, VB_OUT_RAW, 0, 0),                               // conv_res1 / attn_proj: mp_sum (+clip)
    VB_VARIANT24(1, 1, 0, VB_OUT_RAW, VB_OUT_NORM_SILU, 0),
    VB_VARIANT24(1, 1, 0, VB_OUT_RAW, VB_OUT_SILU, 0),
    VB_VARIANT24(1, 1, 0, VB_OUT_RAW, VB_OUT_NORM_SILU, VB_OUT_SILU),
    VB_VARIANT24(1, 1, 0, VB_OUT_RAW, VB_OUT_SILU, VB_OUT_SILU),
    VB_VARIANT24(1, 3, 0, VB_OUT_RAW, 0, 0),                               // ... with the residual scaled per pixel (fused pixel-norm;
    VB_VARIANT24(1, 3, 0, VB_OUT_RAW, VB_OUT_NORM_SILU, 0),                //     VB_RES_PIXNORM itself runs the generic epilogue)
    VB_VARIANT24(1, 3, 0, VB_OUT_RAW, VB_OUT_SILU, 0),
    VB_VARIANT24(1, 3, 0, VB_OUT_RAW, VB_OUT_NORM_SILU, VB_OUT_SILU),
    VB_VARIANT24(1, 3, 0, VB_OUT_RAW, VB_OUT_SILU, VB_OUT_SILU),
    VB_VARIANT(1, -1, -1, -1, -1, -1, 2),                                  // generic staged epilogue (must stay last)
};
#undef VB_VARIANT24
#undef VB_VARIANT_KS
#undef VB_VARIANT

static Variant* find_variant(int staged, int res, int mod, const int* kinds, int parts, int ks = 0) {
  static const bool generic_only = getenv("VB_GENERIC_EPI") != nullptr;       // A/B testing
  if (ks) {                                           // K-split: exact matches only
    for (Variant& v : g_variants)
      if (v.ks && v.staged == staged && v.res == res && v.mod == mod && v.k0 == kinds[0] && v.k1 == kinds[1] && v.k2 == kinds[2]) return &v;
    return nullptr;
  }
  for (int pass = 0; pass < 2; ++pass) {              // second pass: the 8-warp form of whatever was asked for
    const int want = pass == 0 ? parts : 2;
    for (Variant& v : g_variants) {
      if (v.ks || v.staged != staged || v.parts != want) continue;
      if (!staged) return &v;
      if (v.res < 0) return &v;
      if (generic_only) continue;
      if (v.res == res && v.mod == mod && v.k0 == kinds[0] && v.k1 == kinds[1] && v.k2 == kinds[2]) return &v;
    }
  }
  return nullptr;
}

struct ConvLaunch {
  CUtensorMap map_a, map_a2, map_w, map_res;
  OutMaps map_out;
  ConvKernelParams p;
  ConvKernelFn fn;
  int grid;
  int threads;
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

// Choose the main-loop shared-memory layout inside `budget` bytes.  Returns false if not even the minimum fits
// (or, with want_good, if only a shallow pipeline would fit — the caller then retries with a smaller epilogue ring).
static bool plan_mainloop(ConvKernelParams& p, int budget, bool want_good) {
  static const int env_forced = getenv("VB_TAP_MODE") ? atoi(getenv("VB_TAP_MODE")) : -1;    // -1 auto, 0 off (A/B testing)
  const int forced = p.tune_tap == 1 ? 0 : (p.tune_tap == 2 ? 2 : env_forced);
  p.tap_mode = 0;
  p.b_resident = 0;
  const int kct = p.kc_a + p.kc_b;
  if (p.ksplit && p.want_res1) return false;          // K-split streams its half of the weights: no resident layouts
  if (p.taps == 1 && p.want_res1) {
    // 1x1 layer with its N tile's weights resident: K x block_n slab + a ring of activation boxes.  The L2 -> shared-memory
    // fill per tile drops from (128 + block_n) x K x 2 bytes to 128 x K x 2 (the fill rate, ~69 B/clk/SM, is what bounds
    // these layers: profiles/r01_tma_rate.txt).  Explicit request only: does not fit -> not a legal layout.
    const int resident = kct * p.b_bytes;
    const int stages = std::min(kMaxStages, (budget - resident) / kAStageBytes);
    if (stages < 3) return false;
    p.b_resident = 1;
    p.num_stages = stages;
    p.stage_bytes = kAStageBytes;
    p.b_off = stages * kAStageBytes;
    return true;
  }
  if (p.taps == 9 && p.bn == 1 && forced != 0) {
    const int mode = p.bh == 1 ? 1 : 2;
    const int rows = mode == 1 ? p.bw + 2 : (p.bh + 2) * p.bw;
    const int a_tx = rows * 128;
    const int a_slot = (a_tx + 1023) / 1024 * 1024;
    const int resident = 9 * kct * p.b_bytes;
    bool ok = false;
    if (p.n_tiles == 1 && !p.ksplit && budget - resident >= 3 * a_slot) {
      p.b_resident = 1;
      p.a_slots = std::min(kMaxStages, (budget - resident) / a_slot);
      p.b_slots = 0;
      ok = true;
    } else {
      p.b_resident = 0;
      for (int as = 4; as >= 2 && !ok; --as) {
        const int bs = std::min(kMaxBSlots, (budget - as * a_slot) / p.b_bytes);
        if (bs >= (as >= 3 ? 6 : 5)) {
          p.a_slots = as;
          p.b_slots = bs;
          ok = as >= 3 || !want_good;
        }
      }
    }
    // mode 2 only pays when the haloed box is clearly smaller than three separate tap boxes and the pipeline stays deep
    if (ok && (mode == 1 || forced == 2 || (p.bh >= 4 && p.block_n <= 128))) {
      p.tap_mode = mode;
      p.a_slot_bytes = a_slot;
      p.a_tx_bytes = a_tx;
      p.win_rows = mode == 1 ? 1 : p.bw;
      p.b_off = p.a_slots * a_slot;
      return true;
    }
    p.tap_mode = 0;
    p.b_resident = 0;
  }
  const int stages = std::min(kMaxStages, budget / p.stage_bytes);
  if (stages < 2 || (want_good && stages < 3)) return false;
  p.num_stages = stages;
  return true;
}

// Quality of a main-loop layout (higher is better): resident weights >> shared haloed boxes >> per-tap stages, then depth.
static int mainloop_score(const ConvKernelParams& p) {
  if (p.tap_mode != 0) return (p.b_resident ? 2000 : 1000) + 10 * std::min(p.a_slots, 5) + std::min(p.b_slots, 12);
  // 1x1 layers have 2-16 K blocks per tile: four stages cover the TMA latency, and what bounds them is the epilogue — a deep
  // residual ring (4 slots) is worth more than a fifth and sixth stage (profiles/r02_conv1_epilogue.txt)
  return (p.b_resident ? 500 : 0) + 10 * std::min(p.num_stages, p.taps == 1 ? 4 : 6);
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
  for (p.tx_shift = 0; (1 << p.tx_shift) < p.tiles_x; ++p.tx_shift) {}
  for (p.ty_shift = 0; (1 << p.ty_shift) < p.tiles_y; ++p.ty_shift) {}
  const int tiles_nb = (d->B + p.bn - 1) / p.bn;
  p.n_tiles = d->cout_pad / d->block_n;
  p.total_tiles = p.tiles_x * p.tiles_y * tiles_nb * p.n_tiles;
  p.taps = d->taps;
  p.kc_a = d->cin_pad / 64;
  p.kc_b = d->cin2_pad / 64;
  p.block_n = d->block_n;
  static const int env_pair = getenv("VB_PAIR") ? atoi(getenv("VB_PAIR")) : -1;      // -1 auto, 0 off, 1 on (A/B testing)
  static const int env_rowroll = getenv("VB_ROWROLL") ? atoi(getenv("VB_ROWROLL")) : 0;          // A/B testing
  static const bool env_epi_pp = getenv("VB_EPI_PP") != nullptr && atoi(getenv("VB_EPI_PP")) != 0;
  const bool want_rowroll = ((d->tune >> 4) & 3) == 1 || (env_rowroll && ((d->tune >> 4) & 3) == 0 && d->taps == 9 && p.bh == 1 &&
                            p.bw == kBlockM && d->cin_pad == 64 && d->cin2_pad == 0 && d->block_n == 64 && d->cout_pad == 64 &&
                            d->H % 16 == 0 && d->epi_mode == VB_EPI_PLAIN);
  // tune bit 8: K-split over the two CTAs of a cluster (see ConvKernelParams::ksplit).  NOT bitwise-neutral (the K sum is formed as
  // two partial sums), so callers must decide it from the layer's geometry alone — never from the batch size or a timing.
  const bool want_ksplit = ((d->tune >> 8) & 1) != 0;
  VB_REQUIRE(!want_ksplit || (d->ks_ws != nullptr && (p.kc_a + p.kc_b) % 2 == 0 && !want_rowroll && (d->tune & 3) != 2 &&
                              !((d->tune >> 6) & 1) && d->epi_mode == VB_EPI_PLAIN),
             "vb_conv: K-split needs ks_ws, an even number of 64-channel blocks, a plain epilogue, and no pair / ping-pong / row-rolling layout");
  p.ksplit = want_ksplit ? 1 : 0;
  p.ks_ws = static_cast<float*>(d->ks_ws);
  const int forced_pair = (want_rowroll || want_ksplit) ? 0 : ((d->tune & 3) == 1 ? 0 : ((d->tune & 3) == 2 ? 1 : env_pair));
  p.tune_tap = want_rowroll ? 0 : (d->tune >> 2) & 3;
  static const int env_res1 = getenv("VB_RES1") ? atoi(getenv("VB_RES1")) : 0;                   // A/B testing
  p.want_res1 = (d->taps == 1 && (((d->tune >> 7) & 1) || env_res1)) ? 1 : 0;
  const int m_tiles = p.tiles_x * p.tiles_y * tiles_nb;
  // each CTA's half of the weight tile must be whole 8-row swizzle groups, and there must be two M tiles to pair up
  const bool pair_possible = d->block_n % 32 == 0 && m_tiles >= 2;
  auto set_pair = [&](int pair) {
    p.pair = pair;
    p.total_q = pair ? (m_tiles + 1) / 2 * p.n_tiles : p.total_tiles;
    p.b_bytes = (pair ? d->block_n / 2 : d->block_n) * 128;
    p.stage_bytes = kAStageBytes + p.b_bytes;
    p.idesc = umma_idesc_op(pair ? 2 * kBlockM : kBlockM, d->block_n);
  };
  set_pair(0);
  p.epi_mode = d->epi_mode;
  p.flags = d->flags;
  p.mod = d->mod;
  p.mod_stride = d->mod_stride;
  p.out_f32 = d->out_f32;
  p.ld_f32 = d->ld_f32;
  p.out_rnorm = d->out_rnorm;
  p.res_rnorm = d->res_rnorm;
  const float t = d->res_t;
  const float inv = 1.0f / sqrtf((1.f - t) * (1.f - t) + t * t);
  p.res_a = (1.f - t) * inv;
  p.res_b = t * inv;
  // the clip of Block.forward (models.py:204-205) and the saturation of the 16-bit stream are one clamp
#ifdef VB_OP_BF16
  const float finite_max = 3.0e38f;
#else
  const float finite_max = 65504.f;
#endif
  p.clamp = (d->flags & VB_F_CLIP) ? std::min(d->clip, finite_max) : finite_max;
  p.inv_sqrt_c = 1.0f / sqrtf(static_cast<float>(d->cout_pad));
  p.res_slots = 2;
  p.dbg = getenv("VB_DBG") ? atoi(getenv("VB_DBG")) : 0;

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

  int kinds[3] = {0, 0, 0};
  bool staged = false;
  if (d->epi_mode == VB_EPI_PLAIN) {
    if (d->flags & VB_F_MODSILU) VB_REQUIRE_L(d->mod != nullptr && d->mod_stride % 4 == 0, "vb_conv: MODSILU needs mod (stride %% 4)");
    bool needs_norm = false;
    int n_direct = 0, n_norm = 0;
    for (int s = 0; s < 3; ++s) {
      if (d->out[s] == nullptr || d->out_kind[s] == VB_OUT_NONE) continue;
      VB_REQUIRE_L(d->out_kind[s] >= VB_OUT_RAW && d->out_kind[s] <= VB_OUT_NORM_SILU, "vb_conv: bad out_kind[%d]", s);
      VB_REQUIRE_L(p.nslots == s, "vb_conv: output slots must be filled in order");
      p.out_kind[p.nslots] = kinds[p.nslots] = d->out_kind[s];
      p.out_scale[p.nslots] = d->out_scale[s] != 0.f ? d->out_scale[s] : 1.0f;
      if (d->out_kind[s] >= VB_OUT_NORM) {
        needs_norm = true;
        ++n_norm;
      } else {
        ++n_direct;
      }
      ++p.nslots;
    }
    staged = p.nslots > 0;
    p.res_mode = d->res_mode;
    VB_REQUIRE_L(d->res_mode >= VB_RES_NONE && d->res_mode <= VB_RES_SCALED, "vb_conv: bad res_mode");
    VB_REQUIRE_L((d->res_mode == VB_RES_SCALED) == (d->res_rnorm != nullptr), "vb_conv: res_rnorm and VB_RES_SCALED must come together");
    VB_REQUIRE_L((d->res_mode != VB_RES_NONE) == (d->res != nullptr), "vb_conv: res and res_mode must come together");
    VB_REQUIRE_L(p.nslots > 0 || d->out_f32 != nullptr, "vb_conv: no output tensor");
    if (staged) {
      VB_REQUIRE_L(d->block_n % 64 == 0, "vb_conv: staged 16-bit outputs need block_n %% 64 == 0 (got %d)", d->block_n);
      if (d->out_f32) VB_REQUIRE_L(d->ld_f32 % 4 == 0, "vb_conv: ld_f32 must keep 16-byte alignment");
    } else {
      VB_REQUIRE_L(d->res_mode == VB_RES_NONE && !(d->flags & (VB_F_MODSILU | VB_F_CLIP)) && d->ld_f32 % 4 == 0,
                   "vb_conv: the fp32-only epilogue is plain (no residual/modulation/clip)");
    }
    VB_REQUIRE_L(d->out_rnorm == nullptr || needs_norm, "vb_conv: out_rnorm needs a NORM output kind");
    VB_REQUIRE_L(!want_ksplit || staged, "vb_conv: K-split needs 16-bit staged outputs");
    if (needs_norm || d->res_mode == VB_RES_PIXNORM)
      VB_REQUIRE_L(p.n_tiles == 1, "vb_conv: pixel-norm fusion needs the whole channel extent in one tile (cout_pad %d, block_n %d)",
                   d->cout_pad, d->block_n);
    if (staged) {
      // Epilogue shared memory: residual ring + two staging regions of gslots sub-tiles.  A deep residual ring hides the
      // L2 latency of the residual loads when the per-chunk work is short; it is the first thing to shrink.
      p.gslots = std::max(n_direct, n_norm);
      const bool has_res = d->res_mode != VB_RES_NONE;
      static const int forced_regions = getenv("VB_STG_REGIONS") ? atoi(getenv("VB_STG_REGIONS")) : 0;   // A/B testing
      // Shared memory is split between the main loop and the epilogue rings.  Candidates, richest epilogue first; the
      // main loop's needs win (a shallow per-tap pipeline behind a fat epilogue ring cost 2x on the SR 256x256 layers),
      // a CTA pair is taken when it upgrades the main loop (its half-size weight tiles fit resident) or for narrow 3x3
      // tiles, where the single-CTA MMA is shared-memory bound.
      const int opts[4][2] = {{4, 3}, {4, 2}, {2, 3}, {2, 2}};      // {residual ring slots, staging regions}
      // tune bit 6: ping-pong epilogue (two groups, each with its own rings): one-chunk tiles only, two residual slots each
      Variant* egv = find_variant(1, d->res_mode, (d->flags & VB_F_MODSILU) ? 1 : 0, kinds, 1);
      const bool eg_ok = egv != nullptr && egv->parts == 1 && d->block_n <= 128 && !want_rowroll && !want_ksplit && d->res_mode != VB_RES_PIXNORM;
      VB_REQUIRE_L(eg_ok || !((d->tune >> 6) & 1), "vb_conv: no ping-pong epilogue for this layer (needs block_n == 64 and a specialised variant)");
      const int gslots_all = p.gslots;
      bool want_eg = (((d->tune >> 6) & 1) || env_epi_pp) && eg_ok;
      int best_score = -1;
      for (int attempt = 0; attempt < 2 && best_score < 0; ++attempt) {
      if (attempt == 1) {
        if (!want_eg || ((d->tune >> 6) & 1)) break;      // an environment-forced ping-pong epilogue that does not fit: plain one
        want_eg = false;
        p.gslots = gslots_all;
      }
      const int eg_mul = want_eg ? 2 : 1;
      if (want_eg) p.gslots = 1;            // one output per staging slot and commit
      int best[2] = {-1, -1};
      ConvKernelParams best_p[2] = {p, p};
      for (int pair = 0; pair <= ((pair_possible && !want_ksplit) ? 1 : 0); ++pair) {
        for (int o = 0; o < 4; ++o) {
          if (forced_regions == 2 && opts[o][1] != 2) continue;
          if (!has_res && opts[o][0] != 4) continue;               // no residual: the ring size is moot
          if (want_eg && has_res && opts[o][0] != 2) continue;
          set_pair(pair);
          p.res_slots = has_res ? opts[o][0] : 2;
          p.stg_regions = opts[o][1];
          if (!plan_mainloop(p, kSmemMax - eg_mul * ((has_res ? p.res_slots * kChunkBytes : 0) + p.stg_regions * p.gslots * kChunkBytes),
                             false))
            continue;
          const int score = 100 * mainloop_score(p) + (has_res && p.res_slots == 4 ? 40 : 0) + (p.stg_regions == 3 ? 10 : 0);
          if (score > best[pair]) {
            best[pair] = score;
            best_p[pair] = p;
          }
        }
      }
      int use_pair = 0;
      if (best[1] >= 0) {
        const ConvKernelParams& a = best_p[0];
        const ConvKernelParams& b = best_p[1];
        const bool upgrade = b.tap_mode != 0 && b.b_resident && !(best[0] >= 0 && a.tap_mode != 0 && a.b_resident);
        const bool narrow = b.tap_mode != 0 && b.b_resident && d->block_n <= 64;
        use_pair = forced_pair >= 0 ? forced_pair : ((upgrade || narrow) ? 1 : 0);
        if (best[0] < 0) use_pair = 1;
      }
      best_score = best[use_pair];
      p = best_p[use_pair];
      p.eg = want_eg ? 1 : 0;
      }
      VB_REQUIRE_L(best_score >= 0, "vb_conv: shared memory budget exceeded");
    } else {
      if (forced_pair == 1 && pair_possible) set_pair(1);
      VB_REQUIRE_L(plan_mainloop(p, kSmemMax, false), "vb_conv: shared memory budget exceeded");
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
      p.part_scale[j] = d->out_scale[j] != 0.f ? d->out_scale[j] : 1.0f;
    }
    p.norm_scale = 1.0f / sqrtf(static_cast<float>(d->head_dim));
    VB_REQUIRE_L(d->part_ld == 0 || d->part_ld == d->head_dim || (d->part_ld == 64 && d->head_dim == 32),
                 "vb_conv: part_ld must be 0, head_dim, or 64 with head_dim 32");
    p.part_ld = d->part_ld > 0 ? d->part_ld : d->head_dim;
    if (forced_pair == 1 && pair_possible) set_pair(1);
    static const bool qkv_stg_off = getenv("VB_QKV_STAGED") != nullptr && atoi(getenv("VB_QKV_STAGED")) == 0;     // A/B testing
    const int qkv_stage_bytes = 2 * 2 * kChunkBytes;         // two column parts x two slots
    if (!qkv_stg_off && d->head_dim == 64 && p.part_ld == 64 && d->H * d->W >= 64 && d->block_n % 64 == 0 &&
        plan_mainloop(p, kSmemMax - qkv_stage_bytes, true)) {
      p.qkv_stg = 1;
      p.qkv_rows = std::min(kBlockM, d->H * d->W);
    } else {
      VB_REQUIRE_L(plan_mainloop(p, kSmemMax, false), "vb_conv: shared memory budget exceeded");
    }
  }
  if (want_rowroll) {
    VB_REQUIRE_L(d->taps == 9 && p.bh == 1 && p.bw == kBlockM && d->cin_pad == 64 && d->cin2_pad == 0 && d->block_n == 64 &&
                     d->cout_pad == 64 && d->H % 16 == 0 && staged && p.tap_mode == 1 && p.b_resident && !p.pair,
                 "vb_conv: the row-rolling layout needs a 3x3, 64 -> 64 channel layer on rows of >= 128 pixels");
    p.rowroll = 1;
    p.strip_rows = d->H % 32 == 0 ? 32 : 16;
    for (p.strip_shift = 0; (1 << p.strip_shift) < p.strip_rows; ++p.strip_shift) {}
    const int chunks = d->H / p.strip_rows;
    p.chunk_mask = chunks - 1;
    for (p.chunk_shift = 0; (1 << p.chunk_shift) < chunks; ++p.chunk_shift) {}
    p.total_q = p.tiles_x * chunks * d->B;
    p.idesc = umma_idesc_op(kBlockM, 192);
    p.idesc128 = umma_idesc_op(kBlockM, 128);
    p.idesc64 = umma_idesc_op(kBlockM, 64);
  }
  const int main_bytes = p.tap_mode != 0 ? p.b_off + (p.b_resident ? 9 * (p.kc_a + p.kc_b) : p.b_slots) * p.b_bytes
                         : (p.b_resident ? p.b_off + (p.kc_a + p.kc_b) * p.b_bytes : p.num_stages * p.stage_bytes);
  p.res_off = main_bytes;
  const int eg_mul2 = p.eg ? 2 : 1;
  p.stg_off = p.res_off + eg_mul2 * (p.res_mode != VB_RES_NONE ? p.res_slots * kChunkBytes : 0);
  const int qkv_stg_bytes = p.qkv_stg ? 2 * 2 * kChunkBytes : 0;

  // tune bit 6: sixteen epilogue warps (the specialised staged variants only; the row-rolling layout keeps eight)
  const int want_parts = p.eg ? 1 : 2;
  Variant* var = find_variant(staged ? 1 : 0, p.res_mode, (d->flags & VB_F_MODSILU) ? 1 : 0, kinds, want_parts, p.ksplit);
  VB_REQUIRE_L(var != nullptr, p.ksplit ? "vb_conv: K-split is built for single-output (RAW) conv_res0 / conv_res1 epilogues only"
                                        : "vb_conv: no kernel variant");
  VB_REQUIRE_L(var->parts == want_parts || !((d->tune >> 6) & 1), "vb_conv: no ping-pong epilogue variant for this output combination");
  VB_REQUIRE_L(var->parts == want_parts, "vb_conv: ping-pong epilogue unavailable for this output combination");
  l->fn = var->fn;
  l->threads = var->parts == 1 ? 384 : 128 + 128 * var->parts;

  // Tensor maps.  Activations / residual / outputs: {C, W, H, N} with a {64, bw, bh, bn} box;
  // weights: {K, cout_pad} with a {64, block_n} box.
  // (tap modes: the operand box carries the halo — {bw+2, 1} or {bw, bh+2} — and serves three taps)
  const int abw = p.tap_mode == 1 ? p.bw + 2 : p.bw;
  const int abh = p.tap_mode == 2 ? p.bh + 2 : p.bh;
  int rc = encode_act_map(&l->map_a, d->x, d->cin_pad, d->W, d->H, d->B, abw, abh, p.bn);
  if (rc != VB_OK) return fail(rc);
  l->map_a2 = l->map_a;
  if (d->cin2_pad > 0) {
    rc = encode_act_map(&l->map_a2, d->x2, d->cin2_pad, d->W, d->H, d->B, abw, abh, p.bn);
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
  if (p.qkv_stg) {
    // destinations [B/seg_div][heads][seq][64]: 2-D {64, rows} maps, one {64, min(128, H*W)} box per image part of a tile
    for (int j = 0; j < p.parts; ++j) {
      const uint64_t rows = static_cast<uint64_t>(d->B / d->seg_div) * p.heads * d->part_seq[j];
      const uint64_t dims[2] = {64, rows};
      const uint64_t strides[1] = {128};
      const uint32_t box[2] = {64, static_cast<uint32_t>(p.qkv_rows)};
      rc = encode_tmap_16(&l->map_out.m[j], d->part_out[j], 2, dims, strides, box);
      if (rc != VB_OK) return fail(rc);
    }
  }
  {
    const uint64_t ktot = static_cast<uint64_t>(d->taps) * (d->cin_pad + d->cin2_pad);
    const uint64_t dims[2] = {ktot, static_cast<uint64_t>(d->cout_pad)};
    const uint64_t strides[1] = {ktot * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(p.pair ? d->block_n / 2 : d->block_n)};
    rc = encode_tmap_16(&l->map_w, d->w, 2, dims, strides, box);
    if (rc != VB_OK) return fail(rc);
  }
#undef VB_REQUIRE_L

  // The kernel contains cta_group::2 instructions, so the driver only accepts it in clusters of two even when the CTAs work
  // independently (pair == 0): the grid is kept even, a surplus CTA finds no work item and exits.
  l->grid = (p.pair || p.ksplit) ? std::min(2 * p.total_q, num_sms() & ~1)
                                 : std::min(((p.rowroll ? p.total_q : p.total_tiles) + 1) & ~1, num_sms() & ~1);
  if (p.tap_mode == 0 && p.b_resident) {
    // resident 1x1 weights: every CTA must keep ONE N tile -> the work-item stride (CTAs, or pairs) is a multiple of n_tiles
    const int unit = p.n_tiles * (p.pair ? 2 : 1);
    const int step = unit % 2 == 0 ? unit : 2 * unit;
    l->grid = l->grid / step * step;
    if (l->grid == 0) {
      set_error("vb_conv: resident 1x1 weights need at least %d work items", step);
      return fail(VB_ERR_INVALID);
    }
  }
  l->smem_bytes = p.stg_off + eg_mul2 * (staged ? p.stg_regions * p.gslots * kChunkBytes : 0) + qkv_stg_bytes + 1024;
  l->flops = 2.0 * d->B * d->H * d->W * static_cast<double>(d->cout_pad) * d->taps * (d->cin_pad + d->cin2_pad);
  const unsigned long long dev_bit = 1ull << current_device();
  if (!(var->attr_done & dev_bit)) {
    cudaError_t e = cudaFuncSetAttribute(var->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax + 1024);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(conv_gemm_kernel) failed: %s", cudaGetErrorString(e));
      return fail(VB_ERR_CUDA);
    }
    var->attr_done |= dev_bit;
  }
  *out = l;
  return VB_OK;
}

int conv_launch(const ConvLaunch* l, cudaStream_t s, bool chained) {
  static const bool nowait_off = getenv("VB_WGT_NOWAIT") != nullptr && atoi(getenv("VB_WGT_NOWAIT")) == 0;     // A/B testing
  ConvKernelParams kp = l->p;
  kp.wgt_nowait = (chained && !nowait_off) ? 1 : 0;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(l->grid);
  cfg.blockDim = dim3(l->threads);
  cfg.dynamicSmemBytes = l->smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  VB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, l->fn, l->map_a, l->map_a2, l->map_w, l->map_res, l->map_out, kp));
  return VB_OK;
}

void conv_free(ConvLaunch* l) { delete l; }
double conv_flops(const ConvLaunch* l) { return l->flops; }

}  // namespace vb

extern "C" int vb_debug_conv_cycles(unsigned long long* out) {
  VB_REQUIRE(out != nullptr, "vb_debug_conv_cycles: null out");
  unsigned long long zero = 0;
  VB_CHECK_CUDA(cudaMemcpyFromSymbol(out, vb::g_conv_cycles, sizeof(*out)));
  VB_CHECK_CUDA(cudaMemcpyToSymbol(vb::g_conv_cycles, &zero, sizeof(zero)));
  return VB_OK;
}
extern "C" int vb_debug_conv_stamps(long long* out8) {
  VB_REQUIRE(out8 != nullptr, "vb_debug_conv_stamps: null out");
  VB_CHECK_CUDA(cudaMemcpyFromSymbol(out8, vb::g_conv_ts, 8 * sizeof(long long)));
  return VB_OK;
}

extern "C" int64_t vb_conv_ksplit_ws_bytes(int32_t B, int32_t H, int32_t W, int32_t cout_pad) {
  if (B <= 0 || H <= 0 || W <= 0 || cout_pad <= 0) return 0;
  const int64_t bw = std::min(W, vb::kBlockM), bh = std::min<int64_t>(H, vb::kBlockM / bw), bn = vb::kBlockM / (bw * bh);
  const int64_t m_tiles = (W / bw) * (H / bh) * ((B + bn - 1) / bn);
  return m_tiles * vb::kBlockM * cout_pad * 4;
}

extern "C" int vb_conv(const vb_conv_desc* d, void* stream) {
  vb::ConvLaunch* l = nullptr;
  int rc = vb::conv_prepare(d, &l);
  if (rc != VB_OK) return rc;
  rc = vb::conv_launch(l, static_cast<cudaStream_t>(stream), false);
  vb::conv_free(l);
  return rc;
}
