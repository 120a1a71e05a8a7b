// Host-side helpers shared by the translation units of libvividb200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/vivid_b200.h"

namespace vb {

void set_error(const char* fmt, ...);

#define VB_CHECK_CUDA(expr)                                                                       \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      vb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);  \
      return VB_ERR_CUDA;                                                                         \
    }                                                                                             \
  } while (0)

#define VB_REQUIRE(cond, ...)                  \
  do {                                         \
    if (!(cond)) {                             \
      vb::set_error(__VA_ARGS__);              \
      return VB_ERR_INVALID;                   \
    }                                          \
  } while (0)

int num_sms();
int current_device();   // ordinal of the calling thread's current device (0..63)

// Launch with programmatic stream serialisation (unless VB_PDL=0): the kernel's CTAs may start while the previous kernel
// of the stream drains; the kernel MUST call pdl_grid_sync() (ptx.cuh) before its first global-memory access.
bool pdl_enabled();
template <typename... Exp, typename... Act>
cudaError_t launch_pdl(void (*fn)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Act&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, fn, static_cast<Exp>(args)...);
}

// Encode a tiled 16-bit (operand format) tensor map with SWIZZLE_128B (inner box = 64 elements = 128 B).
// dims/strides innermost first; strides in BYTES for dims 1..rank-1.
int encode_tmap_16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box);

// ---- prepared launches (tensor maps encoded once, replayable) ----
struct ConvLaunch;   // conv_gemm.cu
int conv_prepare(const vb_conv_desc* d, ConvLaunch** out);
// chained: the previous kernel of the stream is another op of the same plan (it cannot have written this op's weights)
int conv_launch(const ConvLaunch* l, cudaStream_t s, bool chained);
void conv_free(ConvLaunch* l);
double conv_flops(const ConvLaunch* l);

int attn_launch(const vb_attn_desc* d, cudaStream_t s);
bool attn_tc_supported(const vb_attn_desc* d);
int attn_tc_launch(const vb_attn_desc* d, cudaStream_t s);
int eltwise_launch(const vb_ew_desc* d, cudaStream_t s);
int embed_launch(const vb_emb_desc* d, cudaStream_t s);
int precond_in_launch(const vb_precond_in_desc* d, cudaStream_t s);
int precond_out_launch(const vb_precond_out_desc* d, cudaStream_t s);
int heun_launch(const vb_heun_desc* d, cudaStream_t s);

}  // namespace vb
