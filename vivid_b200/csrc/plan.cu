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
