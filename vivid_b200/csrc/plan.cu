// Plans: a recorded sequence of kernel launches over fixed device buffers, replayed per
// denoiser call either eagerly or as one CUDA graph.  The reference drives the same work
// through ~2.9k eager PyTorch launches per vivid-base call; here the host
// walks the module tree once, records a few hundred fused ops, and replays them.
#include <vector>

#include "common.h"

namespace {

enum OpKind { OP_CONV, OP_ATTN, OP_EW, OP_EMB, OP_PIN, OP_POUT, OP_HEUN };

struct Op {
  OpKind kind;
  vb::ConvLaunch* conv = nullptr;
  union {
    vb_attn_desc attn;
    vb_ew_desc ew;
    vb_emb_desc emb;
    vb_precond_in_desc pin;
    vb_precond_out_desc pout;
    vb_heun_desc heun;
  };
  Op() { memset(&emb, 0, sizeof(emb)); }
};

}  // namespace

struct vb_plan {
  std::vector<Op> ops;
  double flops = 0.0;
  int launches = 0;
  // one instantiated graph per replayed op range: the whole plan, and -- for no_time_enc feature caching
  // (generate_images.py:52-57) -- the source-view encoder and the denoising UNet on their own
  struct Range {
    int first, last;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
  };
  std::vector<Range> graphs;
  vb_io_desc io;
  bool io_bound = false;
};

// chained: op i-1 of the same plan was launched into the stream right before (see vb::conv_launch)
static int run_op(const Op& op, cudaStream_t s, bool chained) {
  switch (op.kind) {
    case OP_CONV: return vb::conv_launch(op.conv, s, chained);
    case OP_ATTN: return vb::attn_launch(&op.attn, s);
    case OP_EW: return vb::eltwise_launch(&op.ew, s);
    case OP_EMB: return vb::embed_launch(&op.emb, s);
    case OP_PIN: return vb::precond_in_launch(&op.pin, s);
    case OP_POUT: return vb::precond_out_launch(&op.pout, s);
    case OP_HEUN: return vb::heun_launch(&op.heun, s);
  }
  return VB_ERR_INVALID;
}

static void drop_graph(vb_plan* p) {
  for (auto& r : p->graphs) {
    if (r.exec) cudaGraphExecDestroy(r.exec);
    if (r.graph) cudaGraphDestroy(r.graph);
  }
  p->graphs.clear();
}

extern "C" int vb_plan_create(vb_plan** out) {
  VB_REQUIRE(out != nullptr, "vb_plan_create: null out");
  *out = new (std::nothrow) vb_plan();
  VB_REQUIRE(*out != nullptr, "vb_plan_create: out of host memory");
  return VB_OK;
}

extern "C" void vb_plan_destroy(vb_plan* p) {
  if (p == nullptr) return;
  drop_graph(p);
  for (Op& op : p->ops)
    if (op.conv) vb::conv_free(op.conv);
  delete p;
}

extern "C" int vb_plan_add_conv(vb_plan* p, const vb_conv_desc* d) {
  VB_REQUIRE(p != nullptr, "vb_plan_add_conv: null plan");
  Op op;
  op.kind = OP_CONV;
  int rc = vb::conv_prepare(d, &op.conv);
  if (rc != VB_OK) return rc;
  p->flops += vb::conv_flops(op.conv);
  p->launches += 1;
  p->ops.push_back(op);
  drop_graph(p);
  return VB_OK;
}

#define VB_PLAN_ADD(NAME, DESC, KIND, FIELD, NLAUNCH)                 \
  extern "C" int NAME(vb_plan* p, const DESC* d) {                    \
    VB_REQUIRE(p != nullptr && d != nullptr, #NAME ": null argument"); \
    Op op;                                                            \
    op.kind = KIND;                                                   \
    op.FIELD = *d;                                                    \
    p->launches += (NLAUNCH);                                         \
    p->ops.push_back(op);                                             \
    drop_graph(p);                                                    \
    return VB_OK;                                                     \
  }

VB_PLAN_ADD(vb_plan_add_eltwise, vb_ew_desc, OP_EW, ew, 1)
VB_PLAN_ADD(vb_plan_add_embed, vb_emb_desc, OP_EMB, emb, (d->mod_total > 0 ? 2 : 1))
VB_PLAN_ADD(vb_plan_add_precond_in, vb_precond_in_desc, OP_PIN, pin, 1)
VB_PLAN_ADD(vb_plan_add_precond_out, vb_precond_out_desc, OP_POUT, pout, 1)
VB_PLAN_ADD(vb_plan_add_heun, vb_heun_desc, OP_HEUN, heun, 1)

extern "C" int vb_plan_add_attn(vb_plan* p, const vb_attn_desc* d) {
  VB_REQUIRE(p != nullptr && d != nullptr, "vb_plan_add_attn: null argument");
  Op op;
  op.kind = OP_ATTN;
  op.attn = *d;
  p->flops += 4.0 * d->B * d->heads * static_cast<double>(d->sq) * d->sk * d->head_dim;
  p->launches += 1;
  p->ops.push_back(op);
  drop_graph(p);
  return VB_OK;
}

extern "C" int vb_plan_num_ops(const vb_plan* p) { return p ? static_cast<int>(p->ops.size()) : 0; }

extern "C" int vb_plan_run(vb_plan* p, int first, int last, void* stream) {
  VB_REQUIRE(p != nullptr, "vb_plan_run: null plan");
  const int n = static_cast<int>(p->ops.size());
  if (last < 0 || last > n) last = n;
  VB_REQUIRE(first >= 0 && first <= last, "vb_plan_run: bad range [%d,%d)", first, last);
  for (int i = first; i < last; ++i) {
    int rc = run_op(p->ops[i], static_cast<cudaStream_t>(stream), i > first);
    if (rc != VB_OK) return rc;
  }
  return VB_OK;
}

extern "C" int vb_plan_launch_graph_range(vb_plan* p, int first, int last, void* stream) {
  VB_REQUIRE(p != nullptr, "vb_plan_launch_graph: null plan");
  const int n = static_cast<int>(p->ops.size());
  if (last < 0 || last > n) last = n;
  VB_REQUIRE(first >= 0 && first < last, "vb_plan_launch_graph: bad range [%d,%d)", first, last);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaGraphExec_t exec = nullptr;
  for (const auto& r : p->graphs)
    if (r.first == first && r.last == last) exec = r.exec;
  if (exec == nullptr) {
    // Capture on a private stream: the caller's stream may be the legacy default stream, which cannot capture.
    cudaStream_t cap = nullptr;
    VB_CHECK_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) {
      cudaStreamDestroy(cap);
      vb::set_error("cudaStreamBeginCapture failed: %s", cudaGetErrorString(e));
      return VB_ERR_CUDA;
    }
    int rc = VB_OK;
    for (int i = first; i < last && rc == VB_OK; ++i) rc = run_op(p->ops[i], cap, i > first);
    cudaGraph_t g = nullptr;
    e = cudaStreamEndCapture(cap, &g);
    cudaStreamDestroy(cap);
    if (rc != VB_OK) {
      if (g) cudaGraphDestroy(g);
      return rc;
    }
    if (e != cudaSuccess) {
      vb::set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
      return VB_ERR_CUDA;
    }
    e = cudaGraphInstantiate(&exec, g, 0);
    if (e != cudaSuccess) {
      cudaGraphDestroy(g);
      vb::set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
      return VB_ERR_CUDA;
    }
    p->graphs.push_back({first, last, g, exec});
  }
  VB_CHECK_CUDA(cudaGraphLaunch(exec, s));
  return VB_OK;
}

extern "C" int vb_plan_launch_graph(vb_plan* p, void* stream) { return vb_plan_launch_graph_range(p, 0, -1, stream); }

extern "C" double vb_plan_query(const vb_plan* p, int kind) {
  if (p == nullptr) return 0.0;
  return kind == 0 ? p->flops : static_cast<double>(p->launches);
}

// ------------------------------------------------------------------------------------------------ whole-call entry point
namespace {
// dst[r][c] = src[(src_rows == 1 ? 0 : r)][c]   (row broadcast of sigma / the pose vector), or zeros when src == nullptr
__global__ void bcast_rows_kernel(float* dst, const float* src, long long rows, long long dim, int src_rows) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * dim) return;
  dst[i] = src == nullptr ? 0.f : src[src_rows == 1 ? i % dim : i];
}
}  // namespace

extern "C" int vb_plan_bind_io(vb_plan* p, const vb_io_desc* io) {
  VB_REQUIRE(p != nullptr && io != nullptr, "vb_plan_bind_io: null argument");
  VB_REQUIRE(io->in_x && io->in_sigma && io->in_geom && io->out_d, "vb_plan_bind_io: in_x, in_sigma, in_geom and out_d are required");
  VB_REQUIRE(io->n_x > 0 && io->n_out > 0 && io->img_elems > 0 && io->geom_dim > 0, "vb_plan_bind_io: bad extents");
  VB_REQUIRE((io->in_cond == nullptr) == (io->in_noise == nullptr), "vb_plan_bind_io: in_cond and in_noise come together");
  p->io = *io;
  p->io_bound = true;
  return VB_OK;
}

extern "C" int64_t vb_workspace_bytes(const vb_plan* p) { return (p != nullptr && p->io_bound) ? p->io.workspace_bytes : 0; }

extern "C" int vb_denoise(vb_plan* p, const float* src, const float* x, const float* sigma, int32_t sigma_n,
                          const float* geometry, int32_t geometry_rows, const float* cond, const float* noise, float* D_out,
                          void* stream) {
  VB_REQUIRE(p != nullptr && p->io_bound, "vb_denoise: the plan has no bound I/O buffers (vb_plan_bind_io)");
  VB_REQUIRE(x != nullptr && sigma != nullptr && D_out != nullptr, "vb_denoise: x, sigma and D_out are required");
  const vb_io_desc& io = p->io;
  VB_REQUIRE(sigma_n == 1 || sigma_n == io.n_x, "vb_denoise: sigma_n must be 1 or %lld (got %d)", static_cast<long long>(io.n_x), sigma_n);
  VB_REQUIRE(geometry == nullptr || geometry_rows == 1 || geometry_rows == io.n_x, "vb_denoise: geometry_rows must be 1 or %lld (got %d)",
             static_cast<long long>(io.n_x), geometry_rows);
  VB_REQUIRE(io.in_src == nullptr || src != nullptr, "vb_denoise: this plan has a source-view encoder: src is required");
  VB_REQUIRE(io.in_cond == nullptr || (cond != nullptr && noise != nullptr), "vb_denoise: super_res plan: cond and noise are required");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t img = static_cast<size_t>(io.img_elems) * sizeof(float);
  VB_CHECK_CUDA(cudaMemcpyAsync(io.in_x, x, io.n_x * img, cudaMemcpyDeviceToDevice, s));
  if (io.in_src != nullptr) VB_CHECK_CUDA(cudaMemcpyAsync(io.in_src, src, io.n_x * img, cudaMemcpyDeviceToDevice, s));
  if (io.in_cond != nullptr) {
    VB_CHECK_CUDA(cudaMemcpyAsync(io.in_cond, cond, io.n_out * img, cudaMemcpyDeviceToDevice, s));
    VB_CHECK_CUDA(cudaMemcpyAsync(io.in_noise, noise, io.n_out * img, cudaMemcpyDeviceToDevice, s));
  }
  bcast_rows_kernel<<<static_cast<unsigned>((io.n_x + 255) / 256), 256, 0, s>>>(io.in_sigma, sigma, io.n_x, 1, sigma_n);
  const long long ng = io.n_x * io.geom_dim;
  bcast_rows_kernel<<<static_cast<unsigned>((ng + 255) / 256), 256, 0, s>>>(io.in_geom, geometry, io.n_x, io.geom_dim, geometry_rows);
  VB_CHECK_CUDA(cudaGetLastError());
  const int rc = vb_plan_launch_graph_range(p, 0, -1, stream);
  if (rc != VB_OK) return rc;
  VB_CHECK_CUDA(cudaMemcpyAsync(D_out, io.out_d, io.n_out * img, cudaMemcpyDeviceToDevice, s));
  return VB_OK;
}
