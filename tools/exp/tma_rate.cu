// Experiment: per-SM throughput of TMA tiled loads (SWIZZLE_128B, 128-byte inner rows) into shared memory as a function
// of box shape, alignment / out-of-bounds start coordinates and the number of loads in flight.  All 148 SMs stream
// through a [32][256][256][64] fp16 tensor (268 MB) the way the conv kernel's activation loads do.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "../../vivid_b200/csrc/ptx.cuh"
using namespace vb;



__global__ void __launch_bounds__(128) tma_kernel(const __grid_constant__ CUtensorMap map, long long* out, int iters, int slots,
                                                  int box_bytes, int xoff, int reuse, int nmask) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[12];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  if (threadIdx.x == 0) { for (int i = 0; i < 12; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x < 32 && elect_one_sync()) {
    const long long t0 = clock64();
    // tight issue loop: no divisions, coordinates advanced incrementally (the first version of this experiment measured
    // its own integer divisions: 555 cycles per iteration whatever the box)
    int s = 0, par = 0, tile = blockIdx.x;
    const uint32_t stride = (box_bytes + 1023) / 1024 * 1024;
    for (int i = 0; i < iters; ++i) {
      if (i >= slots) mbar_wait(&bar[s], par ^ 1);
      mbar_expect_tx(&bar[s], box_bytes);
      tma_load_4d(&map, &bar[s], smem + s * stride, 0, (tile & 1) * 128 + xoff, (tile >> 1) & 255, (tile >> 9) & nmask);
      tile += gridDim.x;
      if (++s == slots) { s = 0; par ^= 1; }
    }
    for (int j = 0; j < slots; ++j) {
      mbar_wait(&bar[s], par ^ 1);
      if (++s == slots) { s = 0; par ^= 1; }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  const size_t elems = 32ull * 256 * 256 * 64;
  __half* d; cudaMalloc(&d, elems * 2); cudaMemset(d, 0, elems * 2);
  long long* dout; cudaMalloc(&dout, 148 * 8);
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int iters = 3000;
  struct Cfg { int bw, bh, xoff, slots, reuse; int promo; int nmask; };
  const Cfg cfgs[] = {
      {128, 1, 0, 8, 1, 0, 31}, {128, 1, 0, 8, 1, 0, 3}, {128, 1, 0, 4, 1, 0, 3}, {128, 1, 0, 2, 1, 0, 3}, {128, 1, 0, 12, 1, 0, 3}, {128, 1, 1, 8, 1, 0, 3},
      {128, 1, -1, 8, 1, 0, 3}, {130, 1, -1, 8, 1, 0, 3}, {130, 1, 0, 8, 1, 0, 3}, {136, 1, 0, 8, 1, 0, 3}, {64, 2, 0, 8, 1, 0, 3},
      {64, 1, 0, 8, 1, 0, 3}, {32, 4, 0, 8, 1, 0, 3}, {32, 1, 0, 8, 1, 0, 3}, {130, 1, -1, 8, 3, 0, 3}, {128, 1, 0, 8, 1, 2, 3}, {130, 1, -1, 12, 1, 2, 3},
      {128, 1, 0, 8, 1, 0, 0}, {256, 1, 0, 6, 1, 0, 3},
  };
  for (const Cfg& c : cfgs) {
    CUtensorMap m;
    cuuint64_t dims[4] = {64, 256, 256, 32}, st[3] = {128, 128 * 256, 128ull * 256 * 256};
    cuuint32_t box[4] = {64, (cuuint32_t)c.bw, (cuuint32_t)c.bh, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     (CUtensorMapL2promotion)c.promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode failed %d\n", r); return 1; }
    const int box_bytes = c.bw * c.bh * 128;
    tma_kernel<<<148, 128, 220 * 1024>>>(m, dout, iters, c.slots, box_bytes, c.xoff, c.reuse, c.nmask);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    long long h[148]; cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("box {64,%3d,%d} x0%+d slots %2d reuse %d l2promo %d nmask %2d : %7.1f cycles per load, %5.1f B/clk/SM\n", c.bw, c.bh, c.xoff, c.slots, c.reuse,
           c.promo, c.nmask, (double)mx / iters, (double)box_bytes * iters / mx);
  }
  return 0;
}
