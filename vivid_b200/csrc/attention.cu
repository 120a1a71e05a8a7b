// Fused cosine attention (flash-style, single pass, no materialised logits).
//
// Replaces einsum -> softmax -> einsum of the snapshot tree
// (experiments/code/training/models.py:190-191, 274-280) and F.scaled_dot_product_attention of
// the current tree (training/models.py:198, 305).  q, k, v arrive already pixel-normalised
// (the qkv GEMM epilogue does it), so |q.k|/sqrt(D) <= sqrt(D): the softmax needs no running
// max — p = exp(q.k/sqrt(D)) stays within [e^-8, e^8], all normal 16-bit values — and no rescaling pass.
// Keys/values of the self segment and of the 1-2 cross (source-view) segments already sit
// back to back in one [B][heads][Sk][D] buffer, so the concat of the reference is free.
// `zero_keys` extra all-zero keys (unconditional gnet) only add exp(0) = 1 each to the denominator.
//
// Data path: cp.async double-buffered K/V tiles in XOR-swizzled shared memory, ldmatrix,
// mma.sync.m16n8k16 (16-bit operands) with fp32 accumulation; 4 warps x 16 query rows per CTA.
// The softmax between the two products is kept to ~1 instruction per logit (v1 spent ~10 and ran at 0.15 of the
// tensor peak, issue-bound): Q is pre-scaled by log2(e)/sqrt(D) once, P = 2^S is ONE packed ex2 on the 16-bit pair
// that is the A fragment of the P.V product anyway (half the MUFU work of an fp32 exp per logit), and the row sums come
// out of the tensor core as P times a ones column instead of per-logit adds.
#include "common.h"
#include "ptx.cuh"

namespace vb {
namespace {

constexpr int kBlockQ = 64;
constexpr int kBlockKV = 64;
constexpr int kAttnThreads = 128;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const uint32_t s = smem_u32(smem);
  const int bytes = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
#ifdef VB_OP_BF16
#define VB_MMA_OP "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32"
#else
#define VB_MMA_OP "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32"
#endif
  asm volatile(
      VB_MMA_OP " {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Packed 16-bit pair helpers for the softmax.
__device__ __forceinline__ uint32_t pack_pair(float lo, float hi) {
#ifdef VB_OP_BF16
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
#else
  __half2 h = __floats2half2_rn(lo, hi);
#endif
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t ex2_pk(uint32_t x) {
  uint32_t r;
#ifdef VB_OP_BF16
  asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(r) : "r"(x));
#else
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(x));
#endif
  return r;
}
__device__ __forceinline__ uint32_t scale_pk(uint32_t x, float c) {
#ifdef VB_OP_BF16
  __nv_bfloat162 y = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&x), __float2bfloat162_rn(c));
#else
  __half2 y = __hmul2(*reinterpret_cast<__half2*>(&x), __float2half2_rn(c));
#endif
  return *reinterpret_cast<uint32_t*>(&y);
}

// Tile of `rows` x D bf16 in smem, 16-byte chunks XOR-swizzled so ldmatrix is conflict-free.
template <int D>
__device__ __forceinline__ int swz(int row, int chunk) {
  constexpr int kChunks = D / 8;
  if (D == 64) return row * kChunks + (chunk ^ (row & 7));
  return row * kChunks + (chunk ^ ((row >> 1) & 3));
}

template <int D>
__device__ __forceinline__ void load_tile_async(op_t* smem, const op_t* gmem, int row0, int rows_total,
                                                int tile_rows) {
  constexpr int kChunks = D / 8;
  for (int i = threadIdx.x; i < tile_rows * kChunks; i += kAttnThreads) {
    const int r = i / kChunks, c = i - r * kChunks;
    const bool ok = row0 + r < rows_total;
    const op_t* src = gmem + static_cast<size_t>(ok ? row0 + r : 0) * D + c * 8;
    cp_async16(smem + swz<D>(r, c) * 8, src, ok);
  }
}

template <int D>
__global__ void __launch_bounds__(kAttnThreads) attn_kernel(const op_t* __restrict__ q,
                                                            const op_t* __restrict__ k,
                                                            const op_t* __restrict__ v,
                                                            op_t* __restrict__ y, int heads, int sq, int sk,
                                                            int zero_keys, int q_prescaled) {
  constexpr int kKSteps = D / 16;    // k-steps of the QK^T product
  constexpr int kDTiles = D / 8;     // n-tiles of the PV product
  __shared__ __align__(128) op_t s_q[kBlockQ * D];
  __shared__ __align__(128) op_t s_k[2][kBlockKV * D];
  __shared__ __align__(128) op_t s_v[2][kBlockKV * D];

  pdl_grid_sync();
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kBlockQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t bh = static_cast<size_t>(b) * heads + h;
  const op_t* qp = q + bh * sq * D;
  const op_t* kp = k + bh * sk * D;
  const op_t* vp = v + bh * sk * D;

  load_tile_async<D>(s_q, qp, q0, sq, kBlockQ);
  load_tile_async<D>(s_k[0], kp, 0, sk, kBlockKV);
  load_tile_async<D>(s_v[0], vp, 0, sk, kBlockKV);
  cp_async_commit();

  const float sqrt_d = sqrtf(static_cast<float>(D));
  // no max subtraction: |logit| <= sqrt(D) <= 8, so exp(logit) in [3e-4, 3e3] is a normal 16-bit value and cannot overflow
  const float c1 = 1.4426950408889634f / sqrt_d;    // log2(e)/sqrt(D), folded into Q
  // B fragment of an 8-column tile whose column 0 is all ones (b[k][n]: lane = 4n + k/2): row sums via the tensor core
#ifdef VB_OP_BF16
  const uint32_t b_ones = lane < 4 ? 0x3F803F80u : 0u;
#else
  const uint32_t b_ones = lane < 4 ? 0x3C003C00u : 0u;
#endif
  const bool ragged = (sk % kBlockKV) != 0;

  uint32_t qf[kKSteps][4];
  float o[kDTiles][4];
#pragma unroll
  for (int j = 0; j < kDTiles; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float lacc[4] = {0.f, 0.f, 0.f, 0.f};   // P x ones: [0] row g, [2] row g+8 (column 0 lives in the lanes with lane%4 == 0)

  const int n_tiles = (sk + kBlockKV - 1) / kBlockKV;
  for (int t = 0; t < n_tiles; ++t) {
    const int cur = t & 1;
    if (t + 1 < n_tiles) {
      load_tile_async<D>(s_k[cur ^ 1], kp, (t + 1) * kBlockKV, sk, kBlockKV);
      load_tile_async<D>(s_v[cur ^ 1], vp, (t + 1) * kBlockKV, sk, kBlockKV);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
      // Q fragments: rows warp*16 + (lane & 15), 16-byte chunk 2*ks + (lane >> 4)
#pragma unroll
      for (int ks = 0; ks < kKSteps; ++ks) {
        const int r = warp * 16 + (lane & 15);
        const int c = 2 * ks + (lane >> 4);
        ldmatrix_x4(smem_u32(s_q + swz<D>(r, c) * 8), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (!q_prescaled) qf[ks][i] = scale_pk(qf[ks][i], c1);       // (else: folded into q by the QKV GEMM epilogue)
      }
    }
    // S = Q K^T : 16 x 64 per warp
    float s[kBlockKV / 8][4];
#pragma unroll
    for (int j = 0; j < kBlockKV / 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < kKSteps; ++ks) {
#pragma unroll
      for (int jp = 0; jp < kBlockKV / 16; ++jp) {
        // four 8x8 blocks: keys jp*16 + {0..7, 8..15}, d chunks 2ks, 2ks+1
        const int r = jp * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int c = 2 * ks + ((lane >> 3) & 1);
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(smem_u32(s_k[cur] + swz<D>(r, c) * 8), b0, b1, b2, b3);
        mma_bf16(s[2 * jp], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
        mma_bf16(s[2 * jp + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
      }
    }
    // P = 2^S on packed 16-bit pairs (S already carries log2(e)/sqrt(D)); keys past the end of the sequence -> 0
    uint32_t pf[kBlockKV / 16][4];
#pragma unroll
    for (int j = 0; j < kBlockKV / 8; ++j) {
      uint32_t p01 = ex2_pk(pack_pair(s[j][0], s[j][1]));
      uint32_t p23 = ex2_pk(pack_pair(s[j][2], s[j][3]));
      if (ragged) {
        const int kk = t * kBlockKV + 2 * (lane & 3) + j * 8;
        const uint32_t m = kk + 1 < sk ? 0xFFFFFFFFu : (kk < sk ? 0x0000FFFFu : 0u);
        p01 &= m;
        p23 &= m;
      }
      pf[j >> 1][(j & 1) * 2 + 0] = p01;
      pf[j >> 1][(j & 1) * 2 + 1] = p23;
    }
    // O += P V, row sums += P 1
#pragma unroll
    for (int kk = 0; kk < kBlockKV / 16; ++kk) {
      mma_bf16(lacc, pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b_ones, b_ones);
#pragma unroll
      for (int dp = 0; dp < kDTiles / 2; ++dp) {
        // V rows (keys) kk*16 + {0..15}, d chunks 2dp, 2dp+1 ; transposed on load
        const int r = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int c = 2 * dp + (lane >> 4);
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(smem_u32(s_v[cur] + swz<D>(r, c) * 8), b0, b1, b2, b3);
        mma_bf16(o[2 * dp], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b0, b1);
        mma_bf16(o[2 * dp + 1], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b2, b3);
      }
    }
    __syncthreads();   // everyone done with buffer `cur` before it is refilled
  }

  // row sums: broadcast from the quad's first lane; plus the analytic zero-key mass (2^0 each)
  const float l0 = __shfl_sync(0xffffffffu, lacc[0], lane & ~3);
  const float l1 = __shfl_sync(0xffffffffu, lacc[2], lane & ~3);
  const float zk = static_cast<float>(zero_keys);
  const float i0 = 1.0f / (l0 + zk), i1 = 1.0f / (l1 + zk);

  // stage the 64 x D output tile through s_q (all Q fragments are in registers by now)
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int j = 0; j < kDTiles; ++j) {
    const int r0 = warp * 16 + g, r1 = r0 + 8;
    const int col = j * 8 + 2 * tq;
    *reinterpret_cast<uint32_t*>(s_q + swz<D>(r0, col >> 3) * 8 + (col & 7)) = pack_op2(o[j][0] * i0, o[j][1] * i0);
    *reinterpret_cast<uint32_t*>(s_q + swz<D>(r1, col >> 3) * 8 + (col & 7)) = pack_op2(o[j][2] * i1, o[j][3] * i1);
  }
  __syncthreads();
  constexpr int kChunks = D / 8;
  const int ldy = heads * D;
  for (int i = threadIdx.x; i < kBlockQ * kChunks; i += kAttnThreads) {
    const int r = i / kChunks, c = i - r * kChunks;
    if (q0 + r < sq) {
      const uint4 val = *reinterpret_cast<const uint4*>(s_q + swz<D>(r, c) * 8);
      *reinterpret_cast<uint4*>(y + (static_cast<size_t>(b) * sq + q0 + r) * ldy + h * D + c * 8) = val;
    }
  }
}

}  // namespace

int attn_launch(const vb_attn_desc* d, cudaStream_t s) {
  VB_REQUIRE(d != nullptr && d->q && d->k && d->v && d->y, "vb_attn: null tensor");
  VB_REQUIRE(d->B > 0 && d->heads > 0 && d->sq > 0 && d->sk > 0, "vb_attn: empty problem");
  VB_REQUIRE(d->head_dim == 64 || d->head_dim == 32, "vb_attn: head_dim must be 32 or 64 (got %d)", d->head_dim);
  VB_REQUIRE(d->zero_keys >= 0, "vb_attn: zero_keys < 0");
  VB_REQUIRE(d->ld == 0 || d->ld == d->head_dim || (d->ld == 64 && d->head_dim == 32), "vb_attn: ld must be 0, head_dim, or 64 with head_dim 32");
  if (attn_tc_supported(d)) return attn_tc_launch(d, s);       // tcgen05 path (attention_tc.cu)
  VB_REQUIRE(d->ld == 0 || d->ld == d->head_dim, "vb_attn: zero-padded rows (ld 64, head_dim 32) need sq %% 256 == 0 and sk %% 128 == 0");
  const dim3 grid((d->sq + kBlockQ - 1) / kBlockQ, d->heads, d->B);
  const op_t* q = static_cast<const op_t*>(d->q);
  const op_t* k = static_cast<const op_t*>(d->k);
  const op_t* v = static_cast<const op_t*>(d->v);
  op_t* y = static_cast<op_t*>(d->y);
  if (d->head_dim == 64)
    VB_CHECK_CUDA(launch_pdl(attn_kernel<64>, grid, dim3(kAttnThreads), 0, s, q, k, v, y, d->heads, d->sq, d->sk, d->zero_keys, d->q_prescaled));
  else
    VB_CHECK_CUDA(launch_pdl(attn_kernel<32>, grid, dim3(kAttnThreads), 0, s, q, k, v, y, d->heads, d->sq, d->sk, d->zero_keys, d->q_prescaled));
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

}  // namespace vb

extern "C" int vb_attn(const vb_attn_desc* d, void* stream) {
  return vb::attn_launch(d, static_cast<cudaStream_t>(stream));
}
