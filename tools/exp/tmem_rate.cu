// Experiment: tcgen05.ld throughput per SM as a function of the number of reading warps and of the load shape.
// One CTA per SM; `nw` warps (warp w reads lane quadrant w%4) each issue `iters` loads of 32 lanes x COLS fp32 columns
// and wait; cycles from clock64 around the loop of the slowest warp.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include "../../vivid_b200/csrc/ptx.cuh"
using namespace vb;

template <int COLS>
__device__ __forceinline__ void ld(uint32_t taddr, float* v) {
  if (COLS == 32) tmem_ld32(taddr, v);
  if (COLS == 16) tmem_ld16(taddr, v);
  if (COLS == 64) { tmem_ld32(taddr, v); tmem_ld32(taddr + 32, v + 32); }
  if (COLS == 128) { tmem_ld32(taddr, v); tmem_ld32(taddr + 32, v + 32); tmem_ld32(taddr + 64, v + 64); tmem_ld32(taddr + 96, v + 96); }
}

template <int COLS>
__global__ void __launch_bounds__(512) rate_kernel(long long* out, float* sink, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    float v[COLS];
    ld<COLS>(tm + ((i * COLS) & 255), v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < COLS; j += 8) acc += v[j];
  }
  const long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

template <int COLS>
void run(long long* d, float* sink, int nw) {
  const int iters = 2000;
  rate_kernel<COLS><<<148, nw * 32>>>(d, sink, iters);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(2); }
  long long h[148 * 16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int b = 0; b < 148; ++b) for (int w = 0; w < nw; ++w) mx = h[b * 16 + w] > mx ? h[b * 16 + w] : mx;
  const double bytes = (double)iters * nw * 32 * COLS * 4;
  printf("cols %3d warps %2d : %.1f cycles per load per warp, %.1f B/clk/SM\n", COLS, nw, (double)mx / iters, bytes / mx);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 16 * 8);
  float* sink; cudaMalloc(&sink, 148 * 512 * 4);
  for (int nw : {1, 2, 4, 8, 16}) { run<16>(d, sink, nw); run<32>(d, sink, nw); run<64>(d, sink, nw); run<128>(d, sink, nw); }
  return 0;
}
