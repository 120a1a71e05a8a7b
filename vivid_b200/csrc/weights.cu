// Weight preparation: forced weight normalisation + magnitude-preserving scaling + gain,
// fp32 statistics, one rounding to the 16-bit operand format, repack for the implicit-GEMM B operand.
// Reference: MPConv.forward prologue, training/models.py:115-121 and normalize :37-42:
//   w = normalize(w.float()) * (gain / sqrt(K))  ==  gain * w / (eps*sqrt(K) + ||w||),  K = cin*taps.
// Runs once per weight version (weights are constant while sampling), not per denoiser call.
#include "common.h"
#include "ptx.cuh"
#include <cuda_fp16.h>

namespace vb {
namespace {

__device__ __forceinline__ float load_w(const void* src, int dtype, size_t i) {
  if (dtype == VB_F32) return static_cast<const float*>(src)[i];
  if (dtype == VB_F16) return __half2float(static_cast<const __half*>(src)[i]);
  return __bfloat162float(static_cast<const __nv_bfloat16*>(src)[i]);
}

// One block per output channel.
__global__ void __launch_bounds__(256) weight_prep_kernel(const vb_weight_prep_desc d) {
  const int co = blockIdx.x;
  const int K = d.cin * d.taps;
  const int cin_pad = d.seg_a_pad + d.seg_b_pad;
  // destination row (qkv de-interleave)
  int row = co;
  if (co < d.cout && d.perm_parts > 0) {
    const int pd = d.perm_parts * d.perm_dim;
    const int h = co / pd, rem = co - h * pd;
    const int dd = rem / d.perm_parts, j = rem - dd * d.perm_parts;
    row = h * pd + j * d.perm_dim + dd;
  }
  if (co >= d.cout) {
    // zero padding rows (16-bit destination only)
    op_t* drow = static_cast<op_t*>(d.dst) + static_cast<size_t>(co) * d.taps * cin_pad;
    for (int i = threadIdx.x; i < d.taps * cin_pad; i += blockDim.x) drow[i] = to_op(0.f);
    return;
  }
  const size_t src0 = static_cast<size_t>(co) * K;
  float ss = 0.f;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const float w = load_w(d.src, d.src_dtype, src0 + i);
    ss += w * w;
  }
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  const float norm = sqrtf(red[0]);
  const float scale = d.gain / (1e-4f * sqrtf(static_cast<float>(K)) + norm);

  if (d.dst_dtype == VB_F32) {
    float* drow = static_cast<float*>(d.dst) + static_cast<size_t>(row) * K;
    for (int i = threadIdx.x; i < K; i += blockDim.x) drow[i] = load_w(d.src, d.src_dtype, src0 + i) * scale;
    return;
  }
  op_t* drow = static_cast<op_t*>(d.dst) + static_cast<size_t>(row) * d.taps * cin_pad;
  for (int i = threadIdx.x; i < d.taps * cin_pad; i += blockDim.x) {
    const int tap = i / cin_pad;
    const int c = i - tap * cin_pad;
    int ci;
    float sc;
    if (c < d.seg_a_pad) {
      ci = c < d.split ? c : -1;
      sc = d.scale_a;
    } else {
      const int cb = c - d.seg_a_pad;
      ci = cb < d.cin - d.split ? d.split + cb : -1;
      sc = d.scale_b;
    }
    float v = 0.f;
    if (ci >= 0) v = load_w(d.src, d.src_dtype, src0 + static_cast<size_t>(ci) * d.taps + tap) * scale * sc;
    drow[i] = to_op(v);
  }
}

}  // namespace
}  // namespace vb

extern "C" int vb_weight_prep(const vb_weight_prep_desc* d, void* stream) {
  VB_REQUIRE(d != nullptr && d->src != nullptr && d->dst != nullptr, "vb_weight_prep: null argument");
  VB_REQUIRE(d->cout > 0 && d->cin > 0 && d->taps > 0, "vb_weight_prep: empty weight");
  VB_REQUIRE(d->src_dtype >= VB_F32 && d->src_dtype <= VB_BF16, "vb_weight_prep: bad src dtype");
  VB_REQUIRE(d->split >= 0 && d->split <= d->cin, "vb_weight_prep: split out of range");
  if (d->dst_dtype == VB_OP_DTYPE) {
    VB_REQUIRE(d->cout_pad >= d->cout, "vb_weight_prep: cout_pad < cout");
    VB_REQUIRE(d->seg_a_pad >= d->split && d->seg_b_pad >= d->cin - d->split, "vb_weight_prep: padded segments too small");
    VB_REQUIRE((d->seg_a_pad + d->seg_b_pad) % 64 == 0, "vb_weight_prep: padded cin must be a multiple of 64");
  } else {
    VB_REQUIRE(d->dst_dtype == VB_F32, "vb_weight_prep: dst dtype must be the operand format (%d) or f32", VB_OP_DTYPE);
    VB_REQUIRE(d->split == d->cin, "vb_weight_prep: fp32 destination has no segments");
  }
  if (d->perm_parts > 0)
    VB_REQUIRE(d->perm_dim > 0 && d->cout % (d->perm_parts * d->perm_dim) == 0, "vb_weight_prep: bad permutation");
  const int rows = d->dst_dtype == VB_OP_DTYPE ? d->cout_pad : d->cout;
  vb::weight_prep_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(*d);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}
