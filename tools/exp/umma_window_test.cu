// Experiment: can a SWIZZLE_128B K-major UMMA A-descriptor start at an arbitrary 128-byte row of a TMA-written tile?
// (needed for the "one haloed row box serves 3 horizontal taps" conv path).  Tries base_offset = 0 and (start>>7)&7.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../../vivid_b200/csrc/ptx.cuh"
using namespace vb;

constexpr int ROWS = 136;   // tile rows loaded by TMA (17 KiB)

__global__ void __launch_bounds__(128) test_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                   float* out, int r0, int use_base_offset) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar, done;
  __shared__ uint32_t slot;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sa = smem;                 // 136 x 128 B
  uint8_t* sb = smem + 18 * 1024;     // 64 x 128 B
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, ROWS * 128 + 64 * 128);
    tma_load_2d(&map_a, &bar, sa, 0, 0);
    tma_load_2d(&map_b, &bar, sb, 0, 0);
    mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_op(128, 64);
    for (int k = 0; k < 4; ++k) {
      const uint32_t start = smem_u32(sa) + r0 * 128 + k * 32;
      uint64_t ad = umma_desc_sw128(start);
      if (use_base_offset) ad |= static_cast<uint64_t>((start >> 7) & 7u) << 49;
      const uint64_t bd = umma_desc_sw128(smem_u32(sb) + k * 32);
      umma_f16_ss(tm, ad, bd, idesc, k != 0);
    }
    umma_commit(&done);
  }
  mbar_wait(&done, 0);
  tc_fence_after();
  float v[32];
  for (int c = 0; c < 64; c += 32) {
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c + j] = v[j];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 64); }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  std::vector<__half> ha(ROWS * 64), hb(64 * 64);
  for (int r = 0; r < ROWS; ++r) for (int c = 0; c < 64; ++c) ha[r * 64 + c] = __float2half((float)(r) + c / 64.0f);
  for (int n = 0; n < 64; ++n) for (int c = 0; c < 64; ++c) hb[n * 64 + c] = __float2half(n == c ? 1.f : 0.f);  // identity: D = A window
  __half *da, *db; float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ma, mb;
  { cuuint64_t d[2] = {64, ROWS}, s[1] = {128}; cuuint32_t b[2] = {64, ROWS}, e[2] = {1, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, da, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode a failed %d\n", r); return 1; } }
  { cuuint64_t d[2] = {64, 64}, s[1] = {128}; cuuint32_t b[2] = {64, 64}, e[2] = {1, 1};
    CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, db, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode b failed %d\n", r); return 1; } }
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  std::vector<float> ho(128 * 64);
  const int r0s[] = {0, 1, 2, 3, 5, 8};
  for (int ubo = 0; ubo < 2; ++ubo)
    for (int r0 : r0s) {
      cudaMemset(dout, 0, 128 * 64 * 4);
      test_kernel<<<1, 128, 40 * 1024>>>(ma, mb, dout, r0, ubo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("r0=%d base_offset=%d: CUDA error %s\n", r0, ubo, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0; float first_bad = 0, first_exp = 0; int bi = -1;
      for (int i = 0; i < 128; ++i) for (int c = 0; c < 64; ++c) {
        const float exp = __half2float(__float2half((float)(r0 + i) + c / 64.0f));
        if (ho[i * 64 + c] != exp) { if (!bad) { first_bad = ho[i * 64 + c]; first_exp = exp; bi = i * 64 + c; } ++bad; }
      }
      printf("r0=%d base_offset_field=%s : %s (%d mismatches; first at %d got %.4f exp %.4f)\n", r0, ubo ? "(start>>7)&7" : "0",
             bad ? "WRONG" : "OK", bad, bi, first_bad, first_exp);
    }
  return 0;
}
