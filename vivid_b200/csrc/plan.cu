// Plans: a recorded sequence of kernel launches over fixed device buffers, replayed per
// denoiser call either eagerly or as one CUDA graph.  The reference drives the same work
// through ~2.9k eager PyTorch launches per vivid-base call; here the host
// walks the module tree once, records a few hundred fused ops, and replays them.
#include <vector>

#include "common.h"
#include "plan.h"

typedef vb_op Op;
typedef vb_op_kind OpKind;

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
  for (void* b : p->owned) cudaFree(b);
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

// ------------------------------------------------------------------------------------------------ whole-sampler entry point
namespace {
// x_hat = noise * t0 (the reference's `noise.to(dtype) * t_steps[0]`, generate_images.py:73), copied into the plans' x inputs;
// the first n_sigma threads also write t0 into their noise-level inputs.
__global__ void sample_init_kernel(const float* noise, float t0, float* x_hat, float* x_in0, float* x_in1, float* sig0, float* sig1,
                                   long long n, long long n_sigma) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = noise[i] * t0;
    x_hat[i] = v;
    x_in0[i] = v;
    if (x_in1 != nullptr) x_in1[i] = v;
  }
  if (i < n_sigma) {
    sig0[i] = t0;
    if (sig1 != nullptr) sig1[i] = t0;
  }
}
}  // namespace

extern "C" int vb_plan_set_inputs(vb_plan* p, const float* src, const float* geometry, int32_t geometry_rows, const float* cond,
                                  void* stream) {
  VB_REQUIRE(p != nullptr && p->io_bound, "vb_plan_set_inputs: the plan has no bound I/O buffers (vb_plan_bind_io)");
  const vb_io_desc& io = p->io;
  VB_REQUIRE(io.in_src == nullptr || src != nullptr, "vb_plan_set_inputs: this plan has a source-view encoder: src is required");
  VB_REQUIRE(io.in_cond == nullptr || cond != nullptr, "vb_plan_set_inputs: super_res plan: cond is required");
  VB_REQUIRE(geometry == nullptr || geometry_rows == 1 || geometry_rows == io.n_x, "vb_plan_set_inputs: geometry_rows must be 1 or %lld (got %d)",
             static_cast<long long>(io.n_x), geometry_rows);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t img = static_cast<size_t>(io.img_elems) * sizeof(float);
  if (io.in_src != nullptr) VB_CHECK_CUDA(cudaMemcpyAsync(io.in_src, src, io.n_x * img, cudaMemcpyDeviceToDevice, s));
  if (io.in_cond != nullptr) VB_CHECK_CUDA(cudaMemcpyAsync(io.in_cond, cond, io.n_out * img, cudaMemcpyDeviceToDevice, s));
  const long long ng = io.n_x * io.geom_dim;
  bcast_rows_kernel<<<static_cast<unsigned>((ng + 255) / 256), 256, 0, s>>>(io.in_geom, geometry, io.n_x, io.geom_dim, geometry_rows);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB_OK;
}

extern "C" int64_t vb_sample_workspace_bytes(const vb_plan* p) {
  return (p != nullptr && p->io_bound) ? 3 * p->io.n_out * p->io.img_elems * static_cast<int64_t>(sizeof(float)) : 0;
}

// The Heun loop of edm_sampler (generate_images.py:72-118; deterministic branch, S_churn = 0) over bound plans: every denoiser
// call is a graph replay whose x / sigma inputs were written by the previous vb_heun pass, the guiding net replays on
// `side_stream` when one is given.  No host synchronisation; the only host work per step is the enqueueing.
extern "C" int vb_sample(const vb_sample_desc* d, void* stream) {
  VB_REQUIRE(d != nullptr && d->net != nullptr && d->net->io_bound, "vb_sample: net must be a plan with bound I/O (vb_plan_bind_io)");
  VB_REQUIRE(d->noise && d->t_steps && d->workspace && d->x_out, "vb_sample: noise, t_steps, workspace and x_out are required");
  VB_REQUIRE(d->num_steps >= 1, "vb_sample: num_steps must be >= 1");
  const bool guided = d->guidance != 1.0f;
  VB_REQUIRE(!guided || (d->gnet != nullptr && d->gnet->io_bound), "vb_sample: guidance != 1 needs a bound gnet plan");
  vb_plan* net = d->net;
  vb_plan* gnet = guided ? d->gnet : nullptr;
  const vb_io_desc& io = net->io;
  VB_REQUIRE(io.n_x == io.n_out, "vb_sample: dual-source plans (2B interleaved inputs) run through the host loop");
  VB_REQUIRE(io.in_noise == nullptr || d->sr_noise != nullptr, "vb_sample: super_res plan: sr_noise (fills in_noise before every call) is required");
  VB_REQUIRE(d->net_first_op >= 0 && d->net_first_op < static_cast<int>(net->ops.size()), "vb_sample: net_first_op out of range");
  if (gnet != nullptr) {
    VB_REQUIRE(gnet != net, "vb_sample: gnet must be a different plan (guidance with gnet == net is the identity: pass guidance = 1)");
    VB_REQUIRE(gnet->io.n_x == io.n_x && gnet->io.n_out == io.n_out && gnet->io.img_elems == io.img_elems && gnet->io.in_noise == nullptr,
               "vb_sample: gnet must be a non-super_res plan of the same batch and resolution");
  }
  for (int i = 0; i < d->num_steps; ++i) VB_REQUIRE(d->t_steps[i] > 0.f, "vb_sample: t_steps[%d] must be > 0", i);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaStream_t side = (gnet != nullptr) ? static_cast<cudaStream_t>(d->side_stream) : nullptr;
  const long long n = io.n_out * io.img_elems;
  float* x_hat = d->workspace;
  float* x_next = d->workspace + n;
  float* d_cur = d->workspace + 2 * n;
  cudaEvent_t ev_in = nullptr, ev_g = nullptr;
  if (side != nullptr) {
    VB_CHECK_CUDA(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
    cudaError_t e = cudaEventCreateWithFlags(&ev_g, cudaEventDisableTiming);
    if (e != cudaSuccess) {
      cudaEventDestroy(ev_in);
      vb::set_error("cudaEventCreateWithFlags failed: %s", cudaGetErrorString(e));
      return VB_ERR_CUDA;
    }
  }
  int rc = VB_OK;
  auto fail_cuda = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && rc == VB_OK) {
      vb::set_error("vb_sample: %s failed: %s", what, cudaGetErrorString(e));
      rc = VB_ERR_CUDA;
    }
  };
  // one denoiser evaluation of both nets; host order net -> gnet as in the reference (generate_images.py:57-62)
  auto denoise = [&]() {
    if (rc != VB_OK) return;
    if (side != nullptr) {
      fail_cuda(cudaEventRecord(ev_in, s), "cudaEventRecord");                 // both nets' inputs are in place
      fail_cuda(cudaStreamWaitEvent(side, ev_in, 0), "cudaStreamWaitEvent");
    }
    if (rc == VB_OK && io.in_noise != nullptr) {
      if (d->sr_noise(d->sr_noise_user, io.in_noise, n, stream) != 0) {
        vb::set_error("vb_sample: the sr_noise callback failed");
        rc = VB_ERR_INVALID;
      }
    }
    if (rc == VB_OK) rc = vb_plan_launch_graph_range(net, d->net_first_op, -1, stream);
    if (rc == VB_OK && gnet != nullptr) {
      rc = vb_plan_launch_graph_range(gnet, 0, -1, side != nullptr ? static_cast<void*>(side) : stream);
      if (rc == VB_OK && side != nullptr) {
        fail_cuda(cudaEventRecord(ev_g, side), "cudaEventRecord");
        fail_cuda(cudaStreamWaitEvent(s, ev_g, 0), "cudaStreamWaitEvent");
      }
    }
  };
  auto heun = [&](int phase, float t_hat, float t_next, bool feed) {
    if (rc != VB_OK) return;
    vb_heun_desc h;
    memset(&h, 0, sizeof(h));
    h.d_net = net->io.out_d;
    h.d_gnet = gnet != nullptr ? gnet->io.out_d : nullptr;
    h.x_hat = x_hat;
    h.d_cur = d_cur;
    h.x_next = x_next;
    h.n = n;
    h.phase = phase;
    h.guidance = d->guidance;
    h.t_hat = t_hat;
    h.t_next = t_next;
    h.sigma_next = t_next;
    if (feed) {                       // x_next and the next call's noise level go straight into the plans' input buffers
      h.x_out[0] = io.in_x;
      h.sigma_out[0] = io.in_sigma;
      if (gnet != nullptr) {
        h.x_out[1] = gnet->io.in_x;
        h.sigma_out[1] = gnet->io.in_sigma;
      }
      h.sigma_n = static_cast<int32_t>(io.n_x);
    }
    rc = vb::heun_launch(&h, s);
  };
  sample_init_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(d->noise, d->t_steps[0], x_hat, io.in_x,
                                                                           gnet != nullptr ? gnet->io.in_x : nullptr, io.in_sigma,
                                                                           gnet != nullptr ? gnet->io.in_sigma : nullptr, n, io.n_x);
  fail_cuda(cudaGetLastError(), "sample_init_kernel");
  for (int i = 0; i < d->num_steps && rc == VB_OK; ++i) {
    const float t_hat = d->t_steps[i], t_next = d->t_steps[i + 1];
    const bool last = i == d->num_steps - 1;
    denoise();
    heun(0, t_hat, t_next, !last);                  // Euler step
    if (!last) {
      denoise();
      heun(1, t_hat, t_next, true);                 // 2nd-order correction
    }
    float* t = x_hat;
    x_hat = x_next;
    x_next = t;
  }
  if (rc == VB_OK) fail_cuda(cudaMemcpyAsync(d->x_out, x_hat, n * sizeof(float), cudaMemcpyDeviceToDevice, s), "cudaMemcpyAsync");
  if (ev_in) cudaEventDestroy(ev_in);       // (destruction is deferred until the recorded work has completed)
  if (ev_g) cudaEventDestroy(ev_g);
  return rc;
}
