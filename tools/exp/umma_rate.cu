// Experiment: issue/execute rate of tcgen05.mma (M=128, kind::f16, SS operands) as a function of N, of how the
// shared-memory descriptors are produced (recomputed from the address vs. advanced by an add), of commit frequency and of
// whether the A window starts at a 1024-byte boundary.  One CTA per SM, one elected thread issues; cycles per MMA from
// clock64 around the whole sequence (after the final commit has completed).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../../vivid_b200/csrc/ptx.cuh"
using namespace vb;

__global__ void __launch_bounds__(128) rate_kernel(long long* out, int N, int iters, int mode, int r0, int commit_every, int coff = 0) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done, dummy[8];
  __shared__ uint32_t slot;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&done, 1); for (int i = 0; i < 8; ++i) mbar_init(&dummy[i], 1); fence_mbar_init(); }
  fence_proxy_async();
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot + coff;
  if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc_op(128, N);
      const uint32_t sa = smem_u32(smem) + r0 * 128;
      const uint32_t sb = smem_u32(smem) + 48 * 1024;
      const long long t0 = clock64();
      if (mode == 0) {
        for (int i = 0; i < iters; ++i) {
          const uint32_t a = sa + (i & 1) * 17408, b = sb + (i & 1) * 16384;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tm, umma_desc_sw128(a + k * 32), umma_desc_sw128(b + k * 32), idesc, (i | k) != 0);
          if (commit_every && (i % commit_every) == commit_every - 1) umma_commit(&dummy[i & 7]);
        }
      } else {
        const uint64_t da0 = umma_desc_sw128(sa), db0 = umma_desc_sw128(sb);
        for (int i = 0; i < iters; ++i) {
          const uint64_t da = da0 + (i & 1) * (17408 >> 4), db = db0 + (i & 1) * (16384 >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tm, da + k * 2, db + k * 2, idesc, (i | k) != 0);
          if (commit_every && (i % commit_every) == commit_every - 1) umma_commit(&dummy[i & 7]);
        }
      }
      umma_commit(&done);
      mbar_wait(&done, 0);
      const long long t1 = clock64();
      out[blockIdx.x] = t1 - t0;
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main(int argc, char** argv) {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  if (argc > 1) {
    // "cal": four launches of back-to-back MMAs (N = 64, 128, 192, 256; the last two keep the tensor pipe 100 % busy) — the
    // calibration run for ncu's tensor-pipe counters (profiles/r02_tensor_pipe_metric.txt)
    for (int N : {64, 128, 192, 256}) {
      rate_kernel<<<148, 128, 100 * 1024>>>(d, N, 4 * iters, 1, 0, 0);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("cal N=%3d : %.1f cycles per MMA (tensor-pipe floor %d), %lld cycles in the MMA sequence\n", N, (double)mx / (iters * 16), N / 2, mx);
    }
    return 0;
  }
  const int Ns[] = {64, 128, 192, 256};
  for (int N : Ns)
    for (int mode = 0; mode < 2; ++mode)
      for (int r0 = 0; r0 < 2; ++r0)
        for (int ce : {0, 1, 3}) {
          rate_kernel<<<148, 128, 100 * 1024>>>(d, N, iters, mode, r0, ce);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
          long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
          long long mx = 0, mn = 1LL << 60; for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
          printf("N=%3d desc=%s window_row=%d commit_every=%d : %.1f .. %.1f cycles per MMA (floor %d)\n", N, mode ? "add " : "calc", r0, ce,
                 (double)mn / (iters * 4), (double)mx / (iters * 4), N / 2);
        }
  // accumulator column offset (the row-rolling conv layout places N=192 accumulators at multiples of 64 columns)
  for (int coff : {0, 64, 128, 192, 256, 320}) {
    rate_kernel<<<148, 128, 100 * 1024>>>(d, 192, iters, 1, 0, 0, coff);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("N=192 accumulator at column %3d : %.1f cycles per MMA\n", coff, (double)mx / (iters * 4));
  }
  return 0;
}
