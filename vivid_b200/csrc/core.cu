// Library core: error reporting, device queries, TMA descriptor encoding.
#include <mutex>

#include <cstdlib>

#include "common.h"

namespace vb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_pdl = -1;      // -1: not decided yet (VB_PDL environment variable), 0 off, 1 on

bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("VB_PDL");
    g_pdl = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}

int current_device() {
  int dev = 0;
  return (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) ? dev : 0;
}

int num_sms() {       // per device ordinal: a process may drive more than one GPU
  static int cached[64] = {0};
  const int dev = current_device();
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached[dev] = n;
    else
      return 148;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#ifdef VB_OP_BF16
static const CUtensorMapDataType kTmapDtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
#else
static const CUtensorMapDataType kTmapDtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
#endif

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_tmap_16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return VB_ERR_NO_DEVICE;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) {
    set_error("tensor map base %p is not 16-byte aligned", base);
    return VB_ERR_INVALID;
  }
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstrides[i] = strides_bytes[i];
  }
  CUresult r = fn(map, kTmapDtype, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdims,
                  gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)", static_cast<int>(r), rank,
              static_cast<unsigned long long>(dims[0]), static_cast<unsigned long long>(dims[1]), box[0], box[1]);
    return VB_ERR_CUDA;
  }
  return VB_OK;
}

// One thread spinning on the global nanosecond timer: holds the stream busy so that launches enqueued behind it run
// back to back (per-op timing in Plan.profile without the host's launch rate in the measurement).
__global__ void spin_kernel(unsigned long long ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < ns);
}

}  // namespace vb

extern "C" int vb_spin(int microseconds, void* stream) {
  VB_REQUIRE(microseconds > 0 && microseconds <= 2000000, "vb_spin: duration out of range");
  vb::spin_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<unsigned long long>(microseconds) * 1000ull);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

extern "C" const char* vb_last_error(void) { return vb::g_err; }
extern "C" int vb_abi_version(void) { return VB_ABI_VERSION; }
#ifdef VB_OP_BF16
extern "C" int vb_operand_dtype(void) { return VB_BF16; }
#else
extern "C" int vb_operand_dtype(void) { return VB_F16; }
#endif
extern "C" int vb_set_pdl(int on) {
  const int prev = vb::pdl_enabled() ? 1 : 0;
  vb::g_pdl = on ? 1 : 0;
  return prev;
}

extern "C" int vb_device_check(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    cudaGetLastError();
    vb::set_error("no CUDA device is available; vivid_b200 has no CPU fallback");
    return VB_ERR_NO_DEVICE;
  }
  if (prop.major != 10) {
    vb::set_error("device %d is sm_%d%d; vivid_b200 kernels are built for sm_100a only", dev, prop.major, prop.minor);
    return VB_ERR_NO_DEVICE;
  }
  return VB_OK;
}

// sizeof of every descriptor struct, so bindings can verify their mirror of the layout.
extern "C" int vb_struct_size(int which) {
  switch (which) {
    case 0: return static_cast<int>(sizeof(vb_weight_prep_desc));
    case 1: return static_cast<int>(sizeof(vb_conv_desc));
    case 2: return static_cast<int>(sizeof(vb_attn_desc));
    case 3: return static_cast<int>(sizeof(vb_ew_desc));
    case 4: return static_cast<int>(sizeof(vb_emb_desc));
    case 5: return static_cast<int>(sizeof(vb_precond_in_desc));
    case 6: return static_cast<int>(sizeof(vb_precond_out_desc));
    case 7: return static_cast<int>(sizeof(vb_heun_desc));
    case 8: return static_cast<int>(sizeof(vb_stats_desc));
    case 9: return static_cast<int>(sizeof(vb_f32_conv_desc));
    case 10: return static_cast<int>(sizeof(vb_f32_op_desc));
    case 11: return static_cast<int>(sizeof(vb_io_desc));
    case 12: return static_cast<int>(sizeof(vb_sample_desc));
    case 13: return static_cast<int>(sizeof(vb_unet_desc));
    case 14: return static_cast<int>(sizeof(vb_net_desc));
    case 15: return static_cast<int>(sizeof(vb_param));
    default: return -1;
  }
}
