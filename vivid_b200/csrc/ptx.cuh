// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences) and UMMA descriptor builders.
// Everything here is device-side plumbing for the kernels in this directory.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace vb {

// The one 16-bit storage format of GEMM operands and the residual stream.  fp16 is the reference's own
// reduced precision (training/models.py:632) and, for these magnitude-preserving networks (activations
// clipped to +-256), 8x more accurate than bf16 at the same tensor-core rate.
#ifdef VB_OP_BF16
typedef __nv_bfloat16 op_t;
#define VB_OP_DTYPE 2
#else
typedef __half op_t;
#define VB_OP_DTYPE 1
#endif

// mbarrier.try_wait suspends the thread until the phase completes OR a time limit passes; with the system's default limit a
// waiting role polls every few dozen nanoseconds (six instructions per failed poll: 15-20 % of all warp-instructions of the
// main-loop-bound conv layers were such polls, profiles/r02_conv_icache.txt).  A longer suspend-time hint keeps the wake-up
// event-driven and removes the polls.  -DVB_WAIT_HINT_NS=0 restores the default limit.
#ifndef VB_WAIT_HINT_NS
#define VB_WAIT_HINT_NS 2000
#endif
#define VB_STR2(x) #x
#define VB_STR(x) VB_STR2(x)
#if VB_WAIT_HINT_NS > 0
#define VB_WAIT_HINT ", " VB_STR(VB_WAIT_HINT_NS)
#else
#define VB_WAIT_HINT ""
#endif
#ifndef VB_SPIN_LIMIT
// a stuck pipeline traps instead of hanging the GPU: with the suspend-time hint a failed poll takes up to VB_WAIT_HINT_NS, so
// 2^23 polls bound a genuine deadlock at ~17 s (without the hint, 2^26 polls of a few dozen nanoseconds each)
#if VB_WAIT_HINT_NS > 0
#define VB_SPIN_LIMIT (1u << 23)
#else
#define VB_SPIN_LIMIT (1u << 26)
#endif
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a fully converged warp.  The compiler recognises elect.sync and issues the uniform-datapath
// instructions (UTCHMMA, UTMALDG, UTCBAR ...) of the guarded region directly; behind a plain `lane == 0` test it
// wraps every one of them in an ELECT/BRA.U.ANY loop (~100 cycles per tcgen05.mma, measured round 1).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Arrive on the barrier at the same smem offset in CTA `cta` of this cluster.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2" VB_WAIT_HINT ";\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > VB_SPIN_LIMIT) {
      printf("vivid_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------- programmatic dependent launch
// Every kernel of the library starts with pdl_grid_sync(): when launched with the programmatic-serialisation attribute
// (common.h launch_pdl) its CTAs may become resident, and run their set-up (barrier init, TMEM allocation, descriptor
// prefetch), while the previous kernel of the stream is still draining; the wait returns once that kernel has completed
// and its writes are visible.  No global memory may be touched before it.  Launched normally both are no-ops.
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// For a role that reads nothing the previous kernel wrote (e.g. a weight producer): no wait, only the release of the next launch.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ----------------------------------------------------------------------------- cluster / CTA pair
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the object at shared::cta address `addr` in CTA `cta` of this cluster.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
// (default .release.cta semantics: a cluster-scope release compiles to MEMBAR.ALL.GPU — ~1.8 k cycles per tile in the
// conv epilogue, measured — and the only data the waiter depends on, TMEM reads, is ordered by tcgen05 fences.)
__device__ __forceinline__ void mbar_arrive_cluster_addr(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}

// ----------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const void* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const void* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// Multicast variants: the box lands at the same smem offset in every CTA of `mask`
// and completes bytes on the mbarrier at the same offset in each of them.
__device__ __forceinline__ void tma_load_2d_mc(const void* map, uint64_t* bar, void* dst, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, "
      "{%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_mc(const void* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                               int c3, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, "
      "{%3, %4, %5, %6}], [%2], %7;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
      : "memory");
}

// CTA-pair (cta_group::2) variants: the box lands in THIS CTA's shared memory, the bytes complete on the mbarrier at
// cluster address `bar_cluster` (the leader CTA's barrier, from mapa_u32) — the pair's MMA issuer waits on one barrier
// for both halves of a stage.
__device__ __forceinline__ void tma_load_2d_pair(const void* map, uint32_t bar_cluster, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(const void* map, uint32_t bar_cluster, void* dst, int c0, int c1, int c2,
                                                 int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// TMA store: shared (SWIZZLE_128B staged tile) -> global, bulk async-group completion.
__device__ __forceinline__ void tma_store_4d(const void* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, one CTA.
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once every tcgen05 op issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// CTA-pair forms: one warp of EACH CTA of the pair allocates / frees; one thread of the leader CTA issues the MMA for
// both (M = 256: rows 0-127 accumulate in the leader's TMEM, 128-255 in the peer's; each CTA supplies its own 128 rows
// of A and its half of B's N rows from the same shared-memory offsets) and commits to the barriers of both.
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_ss_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// Low / high words of the SWIZZLE_128B K-major descriptor (see umma_desc_sw128): the high word is a constant and the
// start-address field of the low word is linear in the byte address (>> 4), so the issuing thread advances descriptors
// with one 32-bit add instead of re-encoding them.
constexpr uint32_t kUmmaDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint64_t umma_desc_from_lo(uint32_t lo) {
  return (static_cast<uint64_t>(kUmmaDescHi) << 32) | lo;
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t <-> lane base+t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// 32 lanes x 32 consecutive 32-bit columns <- the same value from every thread (used to clear an accumulator slot).
__device__ __forceinline__ void tmem_st32_fill(uint32_t taddr, uint32_t v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(v)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile in the canonical SWIZZLE_128B layout: rows of 64 bf16 (128 B),
// groups of 8 rows 1024 B apart (SBO), LBO unused (=1).  cute/atom/mma_traits_sm100.hpp
// documents the same canonical form; this is our own encoder for it.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);   // start address
  d |= static_cast<uint64_t>(1) << 16;                      // leading byte offset (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // stride byte offset: 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;                      // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: 16-bit A/B (both K-major), fp32 accumulate, M x N tile.
__host__ __device__ __forceinline__ uint32_t umma_idesc_op(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                                // C format: F32
#ifdef VB_OP_BF16
  d |= 1u << 7;                                // A format: BF16
  d |= 1u << 10;                               // B format: BF16
#endif                                         // (F16 = 0)
  d |= static_cast<uint32_t>(N >> 3) << 17;    // N / 8
  d |= static_cast<uint32_t>(M >> 4) << 24;    // M / 16
  return d;
}

// ----------------------------------------------------------------------------- math helpers
__device__ __forceinline__ float mp_silu_f(float x) {
  // silu(x) / 0.596  (reference: training/models.py:66-67)
  return __fdividef(x, 1.0f + __expf(-x)) * (1.0f / 0.596f);
}
// silu via one MUFU op: x*sigmoid(x) = x*(0.5 + 0.5*tanh(x/2)); tanh.approx.f32 has ~2^-11 relative error, below
// the 16-bit rounding of the stored result.  Used in the GEMM epilogues where MUFU throughput matters.
__device__ __forceinline__ float mp_silu_fast(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return x * fmaf(t, 0.5f / 0.596f, 0.5f / 0.596f);
}
// Packed fp32 arithmetic (sm_100: FMUL2 / FFMA2 do two lanes' worth of fp32 work per instruction — same results as the scalar
// forms, half the issue slots; the GEMM epilogues are issue- and energy-bound, profiles/r02_energy_note.txt).
__device__ __forceinline__ void mul2(float& a, float& b, float s0, float s1) {          // a *= s0, b *= s1
  asm("{\n.reg .b64 x, s, r;\nmov.b64 x, {%0, %1};\nmov.b64 s, {%2, %3};\nmul.rn.f32x2 r, x, s;\nmov.b64 {%0, %1}, r;\n}"
      : "+f"(a), "+f"(b)
      : "f"(s0), "f"(s1));
}
__device__ __forceinline__ void fma2(float& c0, float& c1, float a0, float a1, float s0, float s1) {   // c = a * s + c
  asm("{\n.reg .b64 x, s, c, r;\nmov.b64 x, {%2, %3};\nmov.b64 s, {%4, %5};\nmov.b64 c, {%0, %1};\nfma.rn.f32x2 r, x, s, c;\n"
      "mov.b64 {%0, %1}, r;\n}"
      : "+f"(c0), "+f"(c1)
      : "f"(a0), "f"(a1), "f"(s0), "f"(s1));
}
// mp_silu of two fp32 values in place: x c (1 + tanh(x/2)), c = 0.5/0.596 (reference training/models.py:66-67)
__device__ __forceinline__ void mp_silu_f32x2(float& a, float& b) {
  float h0 = a, h1 = b;
  mul2(h0, h1, 0.5f, 0.5f);
  float t0, t1;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
  mul2(a, b, 0.5f / 0.596f, 0.5f / 0.596f);
  fma2(a, b, a, b, t0, t1);
}
__device__ __forceinline__ uint32_t pack_op2(float lo, float hi) {
#ifdef VB_OP_BF16
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
#else
  // saturate instead of overflowing to inf (|x| <= 65504; activations are clipped to +-256 by the model): ONE instruction
  // per pair — the explicit fminf/fmaxf form cost four more, a third of the qkv epilogue's arithmetic
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
#endif
#ifdef VB_OP_BF16
  return *reinterpret_cast<uint32_t*>(&h);
#endif
}
__device__ __forceinline__ float2 unpack_op2(uint32_t u) {
#ifdef VB_OP_BF16
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
#else
  return __half22float2(*reinterpret_cast<__half2*>(&u));
#endif
}
__device__ __forceinline__ op_t to_op(float v) {
#ifdef VB_OP_BF16
  return __float2bfloat16(v);
#else
  return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
#endif
}
__device__ __forceinline__ float from_op(op_t v) {
#ifdef VB_OP_BF16
  return __bfloat162float(v);
#else
  return __half2float(v);
#endif
}

}  // namespace vb
