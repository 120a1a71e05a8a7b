// Plan recording inside the library: from a description of the reference's constructor arguments and a table of its
// parameters to a replayable plan with bound I/O (vb_plan_create(net_desc) + vb_plan_set_weights).
//
// What is walked here is the reference's topology — UNet / XAttnUNet / SRXAttnUNet / UNetEncoder constructors
// (training/models.py:340-383, 438-480, 523-534, 575-582), Block / XAttnBlock.forward (:165-206, 251-315), UNet.forward
// (:385-406) and NVPrecond.forward (:628-689; snapshot experiments/code/training/models.py:581-638) — emitted as the fused ops
// of this library (include/vivid_b200.h).  It is the C++ twin of vivid_b200/engine.py and records the same op sequence over the
// same buffer-reuse pattern; `vb_net_plan_trace` / `vb_trace_desc` exist so that a CPU test can compare the two op for op.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.h"
#include "plan.h"

extern "C" int vb_spin(int microseconds, void* stream);

namespace {

// ---------------------------------------------------------------------------------------------------------- trace formatting
// Canonical addresses of a dry run: buffer k lives at (k + 1) << 36, parameter i at (0x4000 + i) << 36.
constexpr int kAddrShift = 36;
constexpr uint64_t kParamBase = 0x4000;

struct Text {
  std::string s;
  void add(const char* fmt, ...) {
    char tmp[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tmp, sizeof(tmp), fmt, ap);
    va_end(ap);
    s += tmp;
  }
  void ptr(const char* key, const void* p) {
    const uint64_t a = reinterpret_cast<uint64_t>(p);
    if (a == 0) {
      add(" %s=-", key);
      return;
    }
    const uint64_t hi = a >> kAddrShift, off = a & ((1ull << kAddrShift) - 1);
    if (hi >= kParamBase) add(" %s=p%llu+%llu", key, static_cast<unsigned long long>(hi - kParamBase), static_cast<unsigned long long>(off));
    else add(" %s=b%llu+%llu", key, static_cast<unsigned long long>(hi - 1), static_cast<unsigned long long>(off));
  }
};

void format_desc(Text& t, int kind, const void* desc) {
  switch (kind) {
    case 0: {
      const auto& d = *static_cast<const vb_weight_prep_desc*>(desc);
      t.add("wprep");
      t.ptr("src", d.src);
      t.ptr("dst", d.dst);
      t.add(" dt=%d>%d cout=%d cin=%d taps=%d cout_pad=%d split=%d seg=%d,%d perm=%d,%d gain=%.9g scale=%.9g,%.9g", d.src_dtype, d.dst_dtype,
            d.cout, d.cin, d.taps, d.cout_pad, d.split, d.seg_a_pad, d.seg_b_pad, d.perm_parts, d.perm_dim, d.gain, d.scale_a, d.scale_b);
      break;
    }
    case 1: {
      const auto& d = *static_cast<const vb_conv_desc*>(desc);
      t.add("conv");
      t.ptr("x", d.x);
      t.ptr("x2", d.x2);
      t.ptr("w", d.w);
      t.ptr("mod", d.mod);
      t.ptr("res", d.res);
      for (int i = 0; i < 3; ++i) t.ptr("out", d.out[i]);
      t.ptr("f32", d.out_f32);
      t.ptr("ornorm", d.out_rnorm);
      t.ptr("rrnorm", d.res_rnorm);
      for (int i = 0; i < 3; ++i) t.ptr("part", d.part_out[i]);
      t.ptr("ks", d.ks_ws);
      t.add(" B=%d H=%d W=%d cin=%d,%d cout_pad=%d taps=%d epi=%d flags=%d mod_stride=%d ld_f32=%d res_mode=%d kinds=%d,%d,%d", d.B, d.H,
            d.W, d.cin_pad, d.cin2_pad, d.cout_pad, d.taps, d.epi_mode, d.flags, d.mod_stride, d.ld_f32, d.res_mode, d.out_kind[0],
            d.out_kind[1], d.out_kind[2]);
      t.add(" D=%d parts=%d seg_div=%d seq=%d,%d,%d off=%d,%d,%d scale=%.9g,%.9g,%.9g res_t=%.9g clip=%.9g part_ld=%d", d.head_dim, d.parts,
            d.seg_div, d.part_seq[0], d.part_seq[1], d.part_seq[2], d.part_off[0], d.part_off[1], d.part_off[2], d.out_scale[0],
            d.out_scale[1], d.out_scale[2], d.res_t, d.clip, d.part_ld);
      // (block_n and tune are the plan-time tuner's, bitwise-neutral: the heuristic start value is part of the trace)
      t.add(" block_n=%d tune=%d", d.block_n, d.tune);
      break;
    }
    case 2: {
      const auto& d = *static_cast<const vb_attn_desc*>(desc);
      t.add("attn");
      t.ptr("q", d.q);
      t.ptr("k", d.k);
      t.ptr("v", d.v);
      t.ptr("y", d.y);
      t.add(" B=%d heads=%d sq=%d sk=%d D=%d zero_keys=%d ld=%d prescaled=%d", d.B, d.heads, d.sq, d.sk, d.head_dim, d.zero_keys, d.ld,
            d.q_prescaled);
      break;
    }
    case 3: {
      const auto& d = *static_cast<const vb_ew_desc*>(desc);
      t.add("eltwise");
      t.ptr("a", d.a);
      t.ptr("b", d.b);
      t.ptr("out", d.out);
      t.ptr("out_silu", d.out_silu);
      t.add(" kind=%d B=%d H=%d W=%d ca=%d cb=%d w=%.9g,%.9g", d.kind, d.B, d.H, d.W, d.ca, d.cb, d.wa, d.wb);
      break;
    }
    case 4: {
      const auto& d = *static_cast<const vb_emb_desc*>(desc);
      t.add("embed");
      t.ptr("sigma", d.sigma);
      t.ptr("geom", d.geom);
      t.ptr("freqs", d.freqs);
      t.ptr("phases", d.phases);
      t.ptr("w_noise", d.w_noise);
      t.ptr("w_label", d.w_label);
      t.ptr("w_mod", d.w_mod);
      t.ptr("emb", d.emb);
      t.ptr("mod", d.mod);
      t.add(" B=%d sigma_n=%d sigma_stride=%d cnoise=%d cemb=%d label_dim=%d mod_total=%d geom_rows=%d balance=%.9g noise_scale=%.9g geom_scale=%.9g",
            d.B, d.sigma_n, d.sigma_stride, d.cnoise, d.cemb, d.label_dim, d.mod_total, d.geom_rows, d.label_balance, d.noise_scale,
            d.geom_scale);
      break;
    }
    case 5: {
      const auto& d = *static_cast<const vb_precond_in_desc*>(desc);
      t.add("precond_in");
      t.ptr("x", d.x);
      t.ptr("cond", d.cond);
      t.ptr("noise", d.noise);
      t.ptr("sigma", d.sigma);
      t.ptr("out", d.out);
      t.add(" B=%d R=%d cpad=%d sigma_n=%d sigma_stride=%d im2col=%d img_stride=%lld sigma_data=%.9g noisy_sr=%.9g", d.B, d.R, d.cpad,
            d.sigma_n, d.sigma_stride, d.im2col, static_cast<long long>(d.img_stride), d.sigma_data, d.noisy_sr);
      break;
    }
    case 6: {
      const auto& d = *static_cast<const vb_precond_out_desc*>(desc);
      t.add("precond_out");
      t.ptr("x", d.x);
      t.ptr("f", d.f);
      t.ptr("sigma", d.sigma);
      t.ptr("d_out", d.d_out);
      t.add(" B=%d R=%d ldf=%d sigma_n=%d sigma_stride=%d img_stride=%lld sigma_data=%.9g", d.B, d.R, d.ldf, d.sigma_n, d.sigma_stride,
            static_cast<long long>(d.img_stride), d.sigma_data);
      break;
    }
    case 7: {
      const auto& d = *static_cast<const vb_io_desc*>(desc);
      t.add("io");
      t.ptr("in_x", d.in_x);
      t.ptr("in_src", d.in_src);
      t.ptr("in_sigma", d.in_sigma);
      t.ptr("in_geom", d.in_geom);
      t.ptr("in_cond", d.in_cond);
      t.ptr("in_noise", d.in_noise);
      t.ptr("out_d", d.out_d);
      t.add(" n_x=%lld n_out=%lld img_elems=%lld geom_dim=%lld workspace_bytes=%lld", static_cast<long long>(d.n_x),
            static_cast<long long>(d.n_out), static_cast<long long>(d.img_elems), static_cast<long long>(d.geom_dim),
            static_cast<long long>(d.workspace_bytes));
      break;
    }
    default: t.add("?");
  }
  t.add("\n");
}

// ---------------------------------------------------------------------------------------------------------- small kernels
// dst[i] = float(src[i])   (the MPFourier buffers of an fp16-persisted net)
__global__ void to_f32_kernel(const void* src, int dtype, float* dst, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dst[i] = dtype == VB_F16 ? __half2float(static_cast<const __half*>(src)[i]) : static_cast<const float*>(src)[i];
}
__global__ void fill_f32_kernel(float* dst, float v, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}

inline int pad_to(int v, int m) { return (v + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------------------- layer table
struct Spec {
  std::string name;      // e.g. "32x32_block1"
  bool enc_group = true; // which ModuleDict holds it
  bool is_conv = false;  // the first MPConv
  int res = 0, cin = 0, cout = 0;
  bool flavor_enc = true;
  int resample = 0;      // 0 keep, 1 up, 2 down
  int heads = 0, head_dim = 0;
  bool xattn = false;
  int skip_ch = 0;
  bool has_conv_skip() const { return !is_conv && cin != cout; }
};

// Layer table of UNet / XAttnUNet (reference training/models.py:340-383 / :438-480); UNetEncoder drops the trailing decoder
// blocks without attention (:523-534).
int unet_layout(const vb_unet_desc& u, std::vector<Spec>& enc, std::vector<Spec>& dec) {
  VB_REQUIRE(u.num_levels >= 1 && u.num_levels <= 8 && u.num_attn_res >= 0 && u.num_attn_res <= 8 && u.num_blocks >= 1 &&
                 u.model_channels > 0 && u.channels_per_head > 0 && u.img_resolution >= 4,
             "vb_net_plan: bad UNet description");
  std::vector<int> widths;
  for (int i = 0; i < u.num_levels; ++i) widths.push_back(u.model_channels * u.channel_mult[i]);
  const int top = u.num_levels - 1;
  auto attention = [&](int res, int slot, int level) {
    for (int i = 0; i < u.num_attn_res; ++i)
      if (u.attn_resolutions[i] == res) return true;
    return u.extra_attn >= 0 && u.extra_attn == slot && level != 0;
  };
  bool bad_heads = false;
  auto mk = [&](const std::string& name, bool enc_group, int cin, int cout, int res, bool flavor_enc, int resample, bool attn) {
    Spec s;
    s.name = name;
    s.enc_group = enc_group;
    s.cin = cin;
    s.cout = cout;
    s.res = res;
    s.flavor_enc = flavor_enc;
    s.resample = resample;
    s.heads = attn ? cout / u.channels_per_head : 0;
    if (attn && s.heads == 0) bad_heads = true;
    s.head_dim = s.heads ? cout / s.heads : 0;
    s.xattn = u.xattn && s.heads > 0;
    return s;
  };
  auto rname = [](int res, const char* what, int idx = -1) {
    std::string n = std::to_string(res) + "x" + std::to_string(res) + "_" + what;
    if (idx >= 0) n += std::to_string(idx);
    return n;
  };
  int width = u.in_channels;
  for (int level = 0; level <= top; ++level) {
    const int ch = widths[level], res = u.img_resolution >> level;
    if (level == 0) {
      Spec s;
      s.name = rname(res, "conv");
      s.is_conv = true;
      s.res = res;
      s.cin = width;
      s.cout = ch;
      enc.push_back(s);
      width = ch;
    } else {
      enc.push_back(mk(rname(res, "down"), true, width, width, res, true, 2, false));
    }
    for (int idx = 0; idx < u.num_blocks; ++idx) {
      enc.push_back(mk(rname(res, "block", idx), true, width, ch, res, true, 0, attention(res, idx, level)));
      width = ch;
    }
  }
  std::vector<int> pending;
  for (const Spec& s : enc) pending.push_back(s.cout);
  for (int level = top; level >= 0; --level) {
    const int ch = widths[level], res = u.img_resolution >> level;
    if (level == top) {
      dec.push_back(mk(rname(res, "in0"), false, width, width, res, false, 0, true));
      dec.push_back(mk(rname(res, "in1"), false, width, width, res, false, 0, false));
    } else {
      dec.push_back(mk(rname(res, "up"), false, width, width, res, false, 1, false));
    }
    for (int idx = 0; idx <= u.num_blocks; ++idx) {
      const int skip = pending.back();
      pending.pop_back();
      Spec b = mk(rname(res, "block", idx), false, width + skip, ch, res, false, 0, attention(res, u.num_blocks - idx, level));
      b.skip_ch = skip;
      dec.push_back(b);
      width = ch;
    }
  }
  VB_REQUIRE(!bad_heads, "vb_net_plan: a block with attention has fewer channels than one head");
  if (u.out_channels == 0)
    while (!dec.empty() && dec.back().heads == 0) dec.pop_back();
  return VB_OK;
}

// ---------------------------------------------------------------------------------------------------------- the recorder
struct Buf {
  void* ptr = nullptr;
  long long numel = 0;      // elements (the activation pool is keyed by element count, like engine.Plan.pool)
  explicit operator bool() const { return ptr != nullptr; }
};

// Block output and the derived forms later consumers read (all 16-bit NHWC [B*R*R, C]); engine.Act.
struct Act {
  int B = 0, R = 0, C = 0;
  Buf raw, norm, nsilu, rnorm;
  std::vector<std::pair<double, Buf>> silu;     // scale -> mp_silu(scale * raw), in insertion order
  bool is_skip = false, is_feature = false;
  Buf silu_of(double scale) const {
    for (const auto& kv : silu)
      if (kv.first == scale) return kv.second;
    return Buf();
  }
  void set_silu(double scale, Buf b) {
    for (auto& kv : silu)
      if (kv.first == scale) {
        kv.second = b;
        return;
      }
    silu.emplace_back(scale, b);
  }
  std::vector<Buf> tensors(bool keep_raw) const {
    std::vector<Buf> out;
    if (!keep_raw) out.push_back(raw);
    out.push_back(norm);
    out.push_back(nsilu);
    out.push_back(rnorm);
    for (const auto& kv : silu) out.push_back(kv.second);
    return out;
  }
};

struct OutSlot {
  Buf t;
  int kind;
  double scale;
};

struct QkvArgs {
  int D = 0, parts = 0, seg_div = 1, ld = 0;
  Buf out[3];
  int seq[3] = {0, 0, 0}, off[3] = {0, 0, 0};
};

struct ConvArgs {
  Buf x2;
  int cin2_pad = 0, cout_pad = 0, flags = 0;
  const float* mod = nullptr;
  int mod_stride = 0;
  Buf res;
  int res_mode = VB_RES_NONE;
  double res_t = 0.3;
  bool has_clip = false;
  double clip = 0.0;
  std::vector<OutSlot> outs;
  Buf out_f32;
  const QkvArgs* qkv = nullptr;
  Buf out_rnorm, res_rnorm;
  bool res_folded = false;
};

typedef std::vector<long long> TuneKey;
std::map<TuneKey, std::pair<int, int>> g_tune_cache;      // process-wide, like engine._TUNE_CACHE; guarded: plans may be recorded
std::mutex g_tune_mutex;                                  // from several host threads (one per device) at once

struct Recorder {
  const vb_net_desc& net;
  const vb_param* params;
  int n_params;
  int B, Bx;
  bool dry;
  cudaStream_t stream;
  vb_plan* plan = nullptr;
  Text trace;
  int op_dtype;
  int n_bufs = 0, n_ops = 0;
  long long owned_bytes = 0;
  std::unordered_map<long long, std::vector<Buf>> pool;
  std::unordered_map<std::string, int> index;
  Buf ks_ws;
  int sm_count = 148;
  bool autotune, fold_res, ksplit;
  int rc = VB_OK;
  // I/O
  Buf in_x, in_src, in_sigma, in_geom, in_cond, in_noise, out_d;
  int geom_dim = 1;

  Recorder(const vb_net_desc& n, const vb_param* p, int np, int batch, bool dry_run, cudaStream_t s)
      : net(n), params(p), n_params(np), B(batch), Bx(n.dual_source ? 2 * batch : batch), dry(dry_run), stream(s) {
    op_dtype = vb_operand_dtype();
    for (int i = 0; i < np; ++i) index[p[i].name] = i;
    const char* e = getenv("VB_AUTOTUNE");
    autotune = !dry && !(e && atoi(e) == 0);
    e = getenv("VB_FOLD_RES");
    fold_res = !(e && atoi(e) == 0);
    e = getenv("VB_KSPLIT");
    ksplit = e && atoi(e) == 1;
    if (!dry) sm_count = vb::num_sms();
  }

  bool ok() const { return rc == VB_OK; }
  int fail(const char* fmt, ...) {
    if (rc == VB_OK) {
      char tmp[512];
      va_list ap;
      va_start(ap, fmt);
      vsnprintf(tmp, sizeof(tmp), fmt, ap);
      va_end(ap);
      vb::set_error("%s", tmp);
      rc = VB_ERR_INVALID;
    }
    return rc;
  }
  void check(int r) {
    if (rc == VB_OK && r != VB_OK) rc = r;
  }
  void check_cuda(cudaError_t e, const char* what) {
    if (rc == VB_OK && e != cudaSuccess) {
      vb::set_error("vb_net_plan: %s failed: %s", what, cudaGetErrorString(e));
      rc = VB_ERR_CUDA;
    }
  }

  // ------------------------------------------------------------------ parameters
  const vb_param* find(const std::string& name) {
    auto it = index.find(name);
    if (it == index.end()) {
      fail("vb_net_plan: parameter '%s' is missing", name.c_str());
      return nullptr;
    }
    return &params[it->second];
  }
  // address of a parameter as the ops see it (dry run: its canonical address)
  const void* param_ptr(const vb_param* p) const {
    if (!dry) return p->data;
    return reinterpret_cast<const void*>((kParamBase + static_cast<uint64_t>(p - params)) << kAddrShift);
  }
  static long long numel_of(const vb_param* p) {
    long long n = 1;
    for (int i = 0; i < p->ndim; ++i) n *= p->shape[i];
    return n;
  }
  // value of a 0-dim parameter (gains); dry run: the table holds HOST pointers
  double scalar(const std::string& name) {
    const vb_param* p = find(name);
    if (p == nullptr) return 0.0;
    if (numel_of(p) != 1 || (p->dtype != VB_F32 && p->dtype != VB_F16)) {
      fail("vb_net_plan: '%s' must be a 0-dim fp32/fp16 parameter", name.c_str());
      return 0.0;
    }
    unsigned char raw[4] = {0, 0, 0, 0};
    const size_t nb = p->dtype == VB_F16 ? 2 : 4;
    if (dry) {
      memcpy(raw, p->data, nb);
    } else {
      check_cuda(cudaMemcpyAsync(raw, p->data, nb, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync");
      check_cuda(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
    }
    if (p->dtype == VB_F16) {
      __half h;
      memcpy(&h, raw, 2);
      return static_cast<double>(__half2float(h));
    }
    float f;
    memcpy(&f, raw, 4);
    return static_cast<double>(f);
  }

  // ------------------------------------------------------------------ buffers
  Buf buf(long long numel, int elem_bytes, bool zero = false) {
    Buf b;
    b.numel = numel;
    const long long bytes = numel * elem_bytes;
    owned_bytes += bytes;
    if (dry) {
      trace.add("alloc b%d %lld z%d\n", n_bufs, bytes, zero ? 1 : 0);
      b.ptr = reinterpret_cast<void*>(static_cast<uint64_t>(n_bufs + 1) << kAddrShift);
    } else if (ok()) {
      void* p = nullptr;
      check_cuda(cudaMalloc(&p, static_cast<size_t>(std::max<long long>(bytes, 16))), "cudaMalloc");
      if (ok()) {
        plan->owned.push_back(p);
        if (zero) check_cuda(cudaMemsetAsync(p, 0, static_cast<size_t>(bytes), stream), "cudaMemsetAsync");
      }
      b.ptr = p;
    }
    ++n_bufs;
    return b;
  }
  // Activations come from a size-keyed pool with LIFO reuse: ops replay in order on one stream, so a buffer whose last consumer
  // has been recorded can back a later activation.
  Buf act(long long rows, int ch) {
    auto it = pool.find(rows * ch);
    if (it != pool.end() && !it->second.empty()) {
      Buf b = it->second.back();
      it->second.pop_back();
      return b;
    }
    return buf(rows * ch, 2);
  }
  Buf a16(int b, int R, int ch) { return act(static_cast<long long>(b) * R * R, ch); }
  void release(const Buf& b) {
    if (b) pool[b.numel].push_back(b);
  }
  void release(const std::vector<Buf>& v) {
    for (const Buf& b : v) release(b);
  }

  // ------------------------------------------------------------------ ops
  void emit(int kind, const void* desc) {
    if (!ok()) return;
    if (dry) {
      format_desc(trace, kind, desc);
      if (kind >= 1 && kind <= 6) ++n_ops;
      return;
    }
    switch (kind) {
      case 0: check(vb_weight_prep(static_cast<const vb_weight_prep_desc*>(desc), stream)); break;
      case 1: check(vb_plan_add_conv(plan, static_cast<const vb_conv_desc*>(desc))); break;
      case 2: check(vb_plan_add_attn(plan, static_cast<const vb_attn_desc*>(desc))); break;
      case 3: check(vb_plan_add_eltwise(plan, static_cast<const vb_ew_desc*>(desc))); break;
      case 4: check(vb_plan_add_embed(plan, static_cast<const vb_emb_desc*>(desc))); break;
      case 5: check(vb_plan_add_precond_in(plan, static_cast<const vb_precond_in_desc*>(desc))); break;
      case 6: check(vb_plan_add_precond_out(plan, static_cast<const vb_precond_out_desc*>(desc))); break;
    }
  }
  int num_ops() const { return dry ? n_ops : (plan ? static_cast<int>(plan->ops.size()) : 0); }

  // vb_weight_prep of parameter `name` viewed as [cout][cin][taps]: normalise (fp32) + gain + pack.  fp32: the plain fp32 matrix
  // (embedding linears), optionally straight into `dst`.
  // gain_param: a 0-dim parameter (emb_gain, out_gain) the constant `gain` is multiplied with.
  Buf prep_weight(const std::string& name, int cout, int cin, int taps, double gain = 1.0, int cout_pad = 0, int perm_parts = 0,
                  int perm_dim = 0, int split = -1, double scale_a = 1.0, double scale_b = 1.0, bool fp32 = false, void* dst = nullptr,
                  const std::string& gain_param = std::string()) {
    const vb_param* p = find(name);
    if (p == nullptr) return Buf();
    const double gain_const = gain;
    if (!gain_param.empty()) gain *= scalar(gain_param);
    if (numel_of(p) != static_cast<long long>(cout) * cin * taps) {
      fail("vb_net_plan: '%s' has %lld elements, the layer table expects %d x %d x %d", name.c_str(), numel_of(p), cout, cin, taps);
      return Buf();
    }
    if (p->dtype != VB_F32 && p->dtype != VB_F16) {
      fail("vb_net_plan: '%s' must be fp32 or fp16", name.c_str());
      return Buf();
    }
    if (split < 0) split = cin;
    vb_weight_prep_desc d;
    memset(&d, 0, sizeof(d));
    d.src = param_ptr(p);
    d.src_dtype = p->dtype;
    d.cout = cout;
    d.cin = cin;
    d.taps = taps;
    d.gain = static_cast<float>(gain);
    Buf out;
    if (fp32) {
      if (dst == nullptr) {
        out = buf(static_cast<long long>(cout) * cin * taps, 4);
        dst = out.ptr;
      }
      d.dst = dst;
      d.dst_dtype = VB_F32;
      d.cout_pad = cout;
      d.split = cin;
      d.seg_a_pad = cin;
      d.seg_b_pad = 0;
      d.scale_a = d.scale_b = 1.0f;
    } else {
      const int sa = pad_to(split, 64), sb = cin > split ? pad_to(cin - split, 64) : 0;
      if (cout_pad == 0) cout_pad = pad_to(cout, 16);
      out = buf(static_cast<long long>(cout_pad) * taps * (sa + sb), 2);
      d.dst = out.ptr;
      d.dst_dtype = op_dtype;
      d.cout_pad = cout_pad;
      d.split = split;
      d.seg_a_pad = sa;
      d.seg_b_pad = sb;
      d.perm_parts = perm_parts;
      d.perm_dim = perm_dim;
      d.scale_a = static_cast<float>(scale_a);
      d.scale_b = static_cast<float>(scale_b);
    }
    emit(0, &d);
    if (!dry && ok()) {
      vb_plan::WeightSlot slot;
      slot.name = name;
      slot.d = d;
      slot.gain_param = gain_param;
      slot.gain_const = gain_const;
      slot.numel = numel_of(p);
      plan->weights.push_back(slot);
    }
    return out;
  }

  // N tile minimising an estimated makespan: waves x (main loop + epilogue) per tile (engine.Plan.pick_block_n).
  int pick_block_n(int cout_pad, long long m_pixels, int k_blocks, int multiple, bool fullrow) const {
    if (fullrow) return cout_pad;
    const long long m_tiles = (m_pixels + 127) / 128;
    int best = 0;
    double best_cost = 0.0;
    const int ns[6] = {256, 192, 128, 64, 32, 16};
    for (int n : ns) {
      if (cout_pad % n || n % multiple) continue;
      const long long tiles = m_tiles * (cout_pad / n);
      const long long waves = (tiles + sm_count - 1) / sm_count;
      const double mma = k_blocks * (std::max(n, 64) / 2.0 + 24);
      const double epi = 6.0 * n + 400;
      const double cost = waves * std::max(mma, epi) + std::min(mma, epi) * 0.15 + 600;
      if (best == 0 || cost < best_cost * 0.999) {
        best = n;
        best_cost = cost;
      }
    }
    return best;
  }

  // Plan-time autotuning of one conv layer (engine.Plan._tune_conv): every legal (block_n, layout) candidate is timed on the
  // layer's own buffers, in batches of launches behind a blocker kernel; the first candidate (the heuristic choice) keeps its
  // place unless another is > 3 % faster.  Candidates differ in tiling only — every output element is computed by the same
  // sequence of MMAs whichever is picked.
  std::pair<int, int> tune_conv(vb_conv_desc d, const std::vector<std::pair<int, int>>& cands) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    check_cuda(cudaEventCreate(&e0), "cudaEventCreate");
    check_cuda(cudaEventCreate(&e1), "cudaEventCreate");
    bool have = false;
    double best_t = 0.0;
    std::pair<int, int> best(d.block_n, d.tune);
    for (const auto& c : cands) {
      if (!ok()) break;
      d.block_n = c.first;
      d.tune = c.second;
      vb_plan* tmp = nullptr;
      if (vb_plan_create(&tmp) != VB_OK) break;
      if (vb_plan_add_conv(tmp, &d) == VB_OK) {
        for (int i = 0; i < 2; ++i) check(vb_plan_run(tmp, 0, -1, stream));
        double t = 1e30;
        int reps = 4;
        for (int batch = 0; batch < 4 && ok(); ++batch) {
          check(vb_spin(batch == 0 ? 80 : 20 * reps, stream));
          check_cuda(cudaEventRecord(e0, stream), "cudaEventRecord");
          for (int i = 0; i < reps; ++i) check(vb_plan_run(tmp, 0, -1, stream));
          check_cuda(cudaEventRecord(e1, stream), "cudaEventRecord");
          check_cuda(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
          float ms = 0.f;
          check_cuda(cudaEventElapsedTime(&ms, e0, e1), "cudaEventElapsedTime");
          if (batch > 0) t = std::min(t, static_cast<double>(ms) / reps);
          else reps = static_cast<int>(std::min(32.0, std::max(4.0, 0.5 / std::max(static_cast<double>(ms) / reps, 1e-3))));
        }
        if (!have || t < best_t * 0.97) {
          have = true;
          best_t = t;
          best = c;
        }
      }
      vb_plan_destroy(tmp);
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (ok() && !have) fail("vb_net_plan: no legal conv layout");
    return best;
  }

  void conv(const Buf& x, const Buf& w, int Bc, int R, int cin_pad, int cout, int taps, const ConvArgs& a) {
    if (!ok()) return;
    const int cout_pad = a.cout_pad ? a.cout_pad : pad_to(cout, 16);
    bool fullrow = a.res_mode == VB_RES_PIXNORM;
    for (const OutSlot& o : a.outs) fullrow = fullrow || o.kind >= VB_OUT_NORM;
    const int multiple = a.qkv ? a.qkv->D : (a.outs.empty() ? 16 : 64);
    const int k_blocks = taps * (cin_pad + a.cin2_pad) / 64;
    int bn = pick_block_n(cout_pad, static_cast<long long>(Bc) * R * R, k_blocks, multiple, fullrow);
    if (bn == 0 || (fullrow && bn > 256)) {
      fail("vb_net_plan: no N tile for a %d-channel layer", cout_pad);
      return;
    }
    int flags = a.flags;
    if (a.has_clip) flags |= VB_F_CLIP;
    if (a.res_folded) flags |= VB_F_RESB_FOLDED;
    vb_conv_desc d;
    memset(&d, 0, sizeof(d));
    d.x = x.ptr;
    d.x2 = a.x2.ptr;
    d.w = w.ptr;
    d.mod = a.mod;
    d.res = a.res.ptr;
    d.out_f32 = static_cast<float*>(a.out_f32.ptr);
    d.out_rnorm = static_cast<float*>(a.out_rnorm.ptr);
    d.res_rnorm = static_cast<const float*>(a.res_rnorm.ptr);
    d.B = Bc;
    d.H = d.W = R;
    d.cin_pad = cin_pad;
    d.cin2_pad = a.cin2_pad;
    d.cout_pad = cout_pad;
    d.taps = taps;
    d.block_n = bn;
    d.epi_mode = a.qkv ? VB_EPI_QKVNORM : VB_EPI_PLAIN;
    d.flags = flags;
    d.mod_stride = a.mod_stride;
    d.ld_f32 = cout_pad;
    d.res_mode = a.res_mode;
    d.res_t = static_cast<float>(a.res_t);
    d.clip = a.has_clip ? static_cast<float>(a.clip) : 0.0f;
    for (size_t i = 0; i < a.outs.size() && i < 3; ++i) {
      d.out[i] = a.outs[i].t.ptr;
      d.out_kind[i] = a.outs[i].kind;
      d.out_scale[i] = static_cast<float>(a.outs[i].scale);
    }
    if (a.qkv) {
      const QkvArgs& q = *a.qkv;
      d.head_dim = q.D;
      d.parts = q.parts;
      d.seg_div = q.seg_div;
      d.part_ld = q.ld;
      // the softmax's log2(e)/sqrt(D) rides on q: one multiply per q element instead of one per logit
      if (q.parts == 3) d.out_scale[0] = static_cast<float>(std::log2(2.718281828459045) / std::sqrt(static_cast<double>(q.D)));
      for (int j = 0; j < q.parts; ++j) {
        d.part_out[j] = q.out[j].ptr;
        d.part_seq[j] = q.seq[j];
        d.part_off[j] = q.off[j];
      }
    }
    const bool modsilu = (flags & VB_F_MODSILU) != 0;
    const bool ks = ksplit && taps == 9 && R <= 8 && ((cin_pad + a.cin2_pad) / 64) % 2 == 0 && !a.qkv && !a.out_f32 && a.outs.size() == 1 &&
                    a.outs[0].kind == VB_OUT_RAW && !a.out_rnorm &&
                    ((a.res_mode == VB_RES_NONE && modsilu) || (a.res_mode == VB_RES_PLAIN && !modsilu));
    if (ks) {
      const long long need = vb_conv_ksplit_ws_bytes(Bc, R, R, cout_pad);
      if (!ks_ws || ks_ws.numel * 4 < need) ks_ws = buf(need / 4, 4);
      d.ks_ws = ks_ws.ptr;
      d.tune = 256;
    }
    if (autotune) {
      TuneKey key = {Bc, R, cin_pad, a.cin2_pad, cout_pad, taps, flags, a.res_mode, a.out_f32 ? 1 : 0, a.mod != nullptr ? 1 : 0, ks ? 1 : 0};
      for (const OutSlot& o : a.outs) key.push_back(o.kind);
      key.push_back(-1);
      if (a.qkv) {
        key.push_back(a.qkv->D);
        key.push_back(a.qkv->parts);
        key.push_back(a.qkv->seg_div);
      }
      key.push_back(vb::current_device());                 // a timing belongs to the device it was taken on
      std::unique_lock<std::mutex> lock(g_tune_mutex);
      auto it = g_tune_cache.find(key);
      std::pair<int, int> choice;
      if (it != g_tune_cache.end()) {
        choice = it->second;
        lock.unlock();
      } else {
        lock.unlock();                                     // (timing runs unlocked: another thread may time the same layer; both results are legal)
        std::vector<int> ns = {bn};
        if (!fullrow)
          for (int n : {256, 192, 128, 64, 32, 16})
            if (n != bn && cout_pad % n == 0 && n % multiple == 0) ns.push_back(n);
        // Only bitwise-neutral knobs are tuned (block_n, single / pair, ping-pong epilogue, resident 1x1 weights): they change
        // the tiling, not the order in which an output element's K terms are summed.
        std::vector<std::pair<int, int>> cands;
        for (int n : ns)
          for (int t = 0; t < 3; ++t) cands.emplace_back(n, t);
        if (!a.outs.empty())
          for (int n : ns)
            if (n <= 128)
              for (int t = 0; t < 3; ++t) cands.emplace_back(n, t | 64);
        if (ks) {
          cands.clear();
          for (int n : ns) cands.emplace_back(n, 256);
        }
        if (taps == 1) {
          for (int n : ns)
            for (int t = 0; t < 3; ++t) cands.emplace_back(n, t | 128);
          if (!a.outs.empty())
            for (int n : ns)
              if (n <= 128)
                for (int t = 0; t < 3; ++t) cands.emplace_back(n, t | 64 | 128);
        }
        choice = tune_conv(d, cands);
        lock.lock();
        g_tune_cache.emplace(key, choice);
        lock.unlock();
      }
      d.block_n = choice.first;
      d.tune = choice.second;
    }
    emit(1, &d);
  }

  void eltwise(int kind, const Buf& a, int Bc, int R, int ca, const Buf& out, const Buf& out_silu) {
    vb_ew_desc d;
    memset(&d, 0, sizeof(d));
    d.a = a.ptr;
    d.out = out.ptr;
    d.out_silu = out_silu.ptr;
    d.kind = kind;
    d.B = Bc;
    d.H = d.W = R;
    d.ca = ca;
    d.wa = d.wb = 1.0f;
    emit(3, &d);
  }

  void attention(const Buf& q, const Buf& k, const Buf& v, const Buf& y, int Bc, int heads, int sq, int sk, int D, int zero_keys, int ld) {
    vb_attn_desc d;
    memset(&d, 0, sizeof(d));
    d.q = q.ptr;
    d.k = k.ptr;
    d.v = v.ptr;
    d.y = y.ptr;
    d.B = Bc;
    d.heads = heads;
    d.sq = sq;
    d.sk = sk;
    d.head_dim = D;
    d.zero_keys = zero_keys;
    d.q_prescaled = 1;       // q carries log2(e)/sqrt(D) already (folded into the QKV GEMM epilogue)
    d.ld = ld;
    emit(2, &d);
  }

  // mp_sum's coefficient of its second operand (training/models.py:71-72)
  static double sum_coeff(double t) { return t / std::sqrt((1.0 - t) * (1.0 - t) + t * t); }

  // ------------------------------------------------------------------ embedding of one UNet (UNet.forward :388-391, Block :175)
  struct Embedded {
    Buf mod;
    std::map<std::string, int> offs;     // "enc/name" | "dec/name" -> first column of the block's modulation vector
    int total = 0;
  };
  static std::string block_key(const Spec& s) { return (s.enc_group ? "enc/" : "dec/") + s.name; }

  Embedded embed(const vb_unet_desc& u, const std::string& prefix, const std::vector<Spec>& specs, int Bc, const Buf& sigma,
                 int sigma_stride, const Buf& geom, int geom_rows, int label_dim, double noise_scale, double geom_scale) {
    Embedded e;
    for (const Spec& s : specs) {
      if (s.is_conv) continue;
      e.offs[block_key(s)] = e.total;
      e.total += s.cout;
    }
    Buf w_mod = buf(static_cast<long long>(e.total) * u.cemb, 4);
    for (const Spec& s : specs) {
      if (s.is_conv) continue;
      const std::string base = prefix + (s.enc_group ? "enc." : "dec.") + s.name + ".";
      prep_weight(base + "emb_linear.weight", s.cout, u.cemb, 1, 1.0, 0, 0, 0, -1, 1.0, 1.0, true,
                  static_cast<char*>(w_mod.ptr) + static_cast<size_t>(e.offs[block_key(s)]) * u.cemb * 4, base + "emb_gain");
    }
    Buf w_noise = prep_weight(prefix + "emb_noise.weight", u.cemb, u.cnoise, 1, 1.0, 0, 0, 0, -1, 1.0, 1.0, true);
    Buf w_label;
    if (u.label_dim != 0) w_label = prep_weight(prefix + "emb_label.weight", u.cemb, u.label_dim, 1, 1.0, 0, 0, 0, -1, 1.0, 1.0, true);
    Buf freqs = buf(u.cnoise, 4), phases = buf(u.cnoise, 4);
    to_f32(freqs, prefix + "emb_fourier.freqs", u.cnoise);
    to_f32(phases, prefix + "emb_fourier.phases", u.cnoise);
    Buf emb = buf(static_cast<long long>(Bc) * u.cemb, 4);
    e.mod = buf(static_cast<long long>(Bc) * e.total, 4);
    vb_emb_desc d;
    memset(&d, 0, sizeof(d));
    d.sigma = static_cast<const float*>(sigma.ptr);
    d.geom = static_cast<const float*>(geom.ptr);
    d.freqs = static_cast<const float*>(freqs.ptr);
    d.phases = static_cast<const float*>(phases.ptr);
    d.w_noise = static_cast<const float*>(w_noise.ptr);
    d.w_label = static_cast<const float*>(w_label.ptr);
    d.w_mod = static_cast<const float*>(w_mod.ptr);
    d.emb = static_cast<float*>(emb.ptr);
    d.mod = static_cast<float*>(e.mod.ptr);
    d.B = Bc;
    d.sigma_n = Bc;
    d.sigma_stride = sigma_stride;
    d.cnoise = u.cnoise;
    d.cemb = u.cemb;
    d.label_dim = label_dim;
    d.mod_total = e.total;
    d.geom_rows = geom_rows;
    d.label_balance = static_cast<float>(u.label_balance);
    d.noise_scale = static_cast<float>(noise_scale);
    d.geom_scale = static_cast<float>(geom_scale);
    emit(4, &d);
    return e;
  }

  void to_f32(const Buf& dst, const std::string& name, long long n) {
    const vb_param* p = find(name);
    if (p == nullptr) return;
    if (numel_of(p) != n || (p->dtype != VB_F32 && p->dtype != VB_F16)) {
      fail("vb_net_plan: buffer '%s' must hold %lld fp32/fp16 values", name.c_str(), n);
      return;
    }
    if (dry) {
      trace.add("to_f32");
      trace.ptr("dst", dst.ptr);
      trace.ptr("src", param_ptr(p));
      trace.add(" n=%lld dt=%d\n", n, p->dtype);
    } else if (ok()) {
      to_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(p->data, p->dtype, static_cast<float*>(dst.ptr), n);
      check_cuda(cudaGetLastError(), "to_f32_kernel");
      vb_plan::WeightSlot slot;
      slot.name = name;
      memset(&slot.d, 0, sizeof(slot.d));
      slot.d.dst = dst.ptr;
      slot.copy = true;
      slot.numel = n;
      plan->weights.push_back(slot);
    }
  }
  void fill(const Buf& dst, float v, long long n) {
    if (dry) {
      trace.add("fill");
      trace.ptr("dst", dst.ptr);
      trace.add(" v=%.9g n=%lld\n", v, n);
    } else if (ok()) {
      fill_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(static_cast<float*>(dst.ptr), v, n);
      check_cuda(cudaGetLastError(), "fill_f32_kernel");
    }
  }

  // mp_cat scale factors (training/models.py:78-84)
  static void cat_weights(int na, int nb, double t, double& wa, double& wb) {
    const double cc = std::sqrt((na + nb) / ((1 - t) * (1 - t) + t * t));
    wa = cc / std::sqrt(static_cast<double>(na)) * (1 - t);
    wb = cc / std::sqrt(static_cast<double>(nb)) * t;
  }

  // ------------------------------------------------------------------ one UNet / encoder (engine.Plan.run_unet)
  // x_in: 16-bit NHWC [B,R,R,64] (im2col of image + ones channels).  features: Acts consumed by cross-attention blocks in order.
  // Returns the raw network output (fp32 [P,16]) for UNets with an out_conv; collected feature Acts go to feats_out.
  Buf run_unet(const vb_unet_desc& u, const std::string& prefix, const std::vector<Spec>& specs, const Buf& x_in, int Bc, const Embedded& e,
               std::vector<Act>* features, int feat_seg, bool zero_feature_keys, bool collect_features, std::vector<Act>* feats_out) {
    const double t_cat = u.concat_balance;
    // mp_cat weights are needed when the PRODUCERS run (they are folded into mp_silu copies / conv_skip weights)
    std::map<std::string, double> skip_scale;
    std::map<std::string, std::pair<double, double>> cat_scale;
    {
      std::vector<const Spec*> pending;
      int width = 0;
      for (const Spec& s : specs) {
        if (s.enc_group) {
          pending.push_back(&s);
          width = s.cout;
        } else {
          if (s.skip_ch) {
            const Spec* en = pending.back();
            pending.pop_back();
            double wa, wb;
            cat_weights(width, en->cout, t_cat, wa, wb);
            skip_scale[en->name] = wb;
            cat_scale[s.name] = std::make_pair(wa, wb);
          }
          width = s.cout;
        }
      }
    }
    struct Form {
      int attr;      // 0 raw, 1 nsilu, 2 silu
      int kind;
      double scale;
      bool operator==(const Form& o) const { return attr == o.attr && kind == o.kind && scale == o.scale; }
    };
    // Output forms block i must emit, in slot order
    auto want = [&](size_t i) {
      const Spec& s = specs[i];
      const Spec* nxt = i + 1 < specs.size() ? &specs[i + 1] : nullptr;
      const bool fullrow = s.cout <= 256;
      std::vector<Form> forms = {{0, VB_OUT_RAW, 1.0}};
      auto push = [&](Form f) {
        if (std::find(forms.begin(), forms.end(), f) == forms.end()) forms.push_back(f);     // e.g. 8x8_block2: silu(1.0) serves both in0 and the first mp_cat
      };
      if (nxt != nullptr && !nxt->is_conv) {
        if (nxt->flavor_enc && nxt->resample == 0 && !nxt->has_conv_skip() && fullrow) push({1, VB_OUT_NORM_SILU, 1.0});
        else if (!nxt->flavor_enc && nxt->resample == 0 && !nxt->skip_ch) push({2, VB_OUT_SILU, 1.0});
        else if (!nxt->flavor_enc && nxt->skip_ch) push({2, VB_OUT_SILU, cat_scale[nxt->name].first});
      }
      if (s.enc_group && skip_scale.count(s.name)) push({2, VB_OUT_SILU, skip_scale[s.name]});
      return forms;
    };
    auto alloc_outs = [&](Act& out, const std::vector<Form>& forms) {
      std::vector<OutSlot> outs;
      for (const Form& f : forms) {
        Buf t = a16(out.B, out.R, out.C);
        if (f.attr == 2) out.set_silu(f.scale, t);
        else if (f.attr == 1) out.nsilu = t;
        else out.raw = t;
        if (f.attr == 1) out.rnorm = act(static_cast<long long>(out.B) * out.R * out.R, 2);   // the consumer's residual scale travels with it
        outs.push_back({t, f.kind, f.scale});
      }
      return outs;
    };

    std::vector<Act> skips;
    Act cur;
    bool have_cur = false;
    size_t next_feature = 0;
    for (size_t i = 0; i < specs.size() && ok(); ++i) {
      const Spec& s = specs[i];
      const std::string base = prefix + (s.enc_group ? "enc." : "dec.") + s.name + ".";
      Act out;
      out.B = Bc;
      out.R = s.res;
      out.C = s.cout;
      out.is_skip = s.enc_group;
      out.is_feature = collect_features && s.heads > 0;
      const int R = s.res, Cc = s.cout;
      std::vector<Buf> temps;
      bool popped = false;
      Act popped_skip;
      Buf res_rnorm;

      if (s.is_conv) {
        // x_in holds the im2col'd 3x3 neighbourhood (vb_precond_in, im2col = 1): the first conv is a K = 64 1x1 GEMM
        Buf w = prep_weight(base + "weight", Cc, s.cin * 9, 1);
        ConvArgs a;
        a.outs = alloc_outs(out, want(i));
        a.out_rnorm = out.rnorm;
        conv(x_in, w, Bc, R, 64, Cc, 1, a);
        cur = out;
        have_cur = true;
        skips.push_back(out);
        continue;
      }
      if (!have_cur) {
        fail("vb_net_plan: the layer table does not start with the input conv");
        break;
      }
      const bool fullrow = Cc <= 256;
      Buf a0, x2, res;
      int k0 = 0, k2 = 0, res_mode = VB_RES_PLAIN;
      // ---------------- main branch: residual base + conv_res0 operand(s)
      if (s.flavor_enc) {
        if (s.resample == 2) {
          Buf bs = a16(Bc, R, Cc);
          a0 = a16(Bc, R, Cc);
          temps.push_back(bs);
          temps.push_back(a0);
          eltwise(VB_EW_DOWN_PIXNORM, cur.raw, Bc, R, Cc, bs, a0);
          res = bs;
        } else if (s.has_conv_skip()) {
          Buf bs = a16(Bc, R, Cc);
          a0 = a16(Bc, R, Cc);
          temps.push_back(bs);
          temps.push_back(a0);
          Buf w = prep_weight(base + "conv_skip.weight", Cc, s.cin, 1);
          ConvArgs a;
          if (fullrow) {           // x = normalize(conv_skip(x)) in one GEMM
            a.outs = {{bs, VB_OUT_NORM, 1.0}, {a0, VB_OUT_NORM_SILU, 1.0}};
            conv(cur.raw, w, Bc, R, pad_to(s.cin, 64), Cc, 1, a);
          } else {
            Buf tmp = a16(Bc, R, Cc);
            temps.push_back(tmp);
            a.outs = {{tmp, VB_OUT_RAW, 1.0}};
            conv(cur.raw, w, Bc, R, pad_to(s.cin, 64), Cc, 1, a);
            eltwise(VB_EW_PIXNORM, tmp, Bc, R, Cc, bs, a0);
          }
          res = bs;
        } else if (cur.nsilu) {     // pixel-norm fused on both sides
          a0 = cur.nsilu;
          res = cur.raw;
          res_mode = VB_RES_SCALED;
          res_rnorm = cur.rnorm;
        } else {
          Buf bs = a16(Bc, R, Cc);
          a0 = a16(Bc, R, Cc);
          temps.push_back(bs);
          temps.push_back(a0);
          eltwise(VB_EW_PIXNORM, cur.raw, Bc, R, Cc, bs, a0);
          res = bs;
        }
        k0 = Cc;
      } else {
        if (s.resample == 1) {
          Buf bs = a16(Bc, R, Cc);
          a0 = a16(Bc, R, Cc);
          temps.push_back(bs);
          temps.push_back(a0);
          eltwise(VB_EW_UP, cur.raw, Bc, R, Cc, bs, a0);
          res = bs;
          k0 = Cc;
        } else if (s.skip_ch) {
          if (skips.empty()) {
            fail("vb_net_plan: decoder block %s has no skip to concatenate", s.name.c_str());
            break;
          }
          popped_skip = skips.back();
          skips.pop_back();
          popped = true;
          const int na = cur.C, nb = popped_skip.C;
          if (nb != s.skip_ch || na + nb != s.cin || na % 64 || nb % 64) {
            fail("vb_net_plan: %s: channel counts %d + %d cannot be folded into a two-source K loop", s.name.c_str(), na, nb);
            break;
          }
          const double wa = cat_scale[s.name].first, wb = cat_scale[s.name].second;
          a0 = cur.silu_of(wa);
          x2 = popped_skip.silu_of(wb);
          k0 = na;
          k2 = nb;
          Buf bs = a16(Bc, R, Cc);
          temps.push_back(bs);
          Buf w = prep_weight(base + "conv_skip.weight", Cc, s.cin, 1, 1.0, 0, 0, 0, na, wa, wb);
          ConvArgs a;
          a.x2 = popped_skip.raw;
          a.cin2_pad = nb;
          a.outs = {{bs, VB_OUT_RAW, 1.0}};
          conv(cur.raw, w, Bc, R, na, Cc, 1, a);
          res = bs;
        } else {
          a0 = cur.silu_of(1.0);
          res = cur.raw;
          k0 = Cc;
        }
      }
      if (k0 % 64) {
        fail("vb_net_plan: %s: %d input channels are not a multiple of 64", s.name.c_str(), k0);
        break;
      }
      if (!a0 || (k2 && !x2)) {
        fail("vb_net_plan: %s: the producer did not emit the activation form this block reads", s.name.c_str());
        break;
      }

      // ---------------- residual branch
      Buf y0 = a16(Bc, R, Cc);
      temps.push_back(y0);
      Buf w0 = prep_weight(base + "conv_res0.weight", Cc, k0 + k2, 9, 1.0, 0, 0, 0, x2 ? k0 : -1);
      {
        ConvArgs a;
        a.x2 = x2;
        a.cin2_pad = k2;
        a.flags = VB_F_MODSILU;
        a.mod = reinterpret_cast<const float*>(static_cast<const char*>(e.mod.ptr) + 4 * static_cast<size_t>(e.offs.at(block_key(s))));
        a.mod_stride = e.total;
        a.outs = {{y0, VB_OUT_RAW, 1.0}};
        conv(a0, w0, Bc, R, k0, Cc, 9, a);
      }
      // mp_sum(x, y, t) = (x (1-t) + y t) / sqrt((1-t)^2 + t^2): y's coefficient rides on the prepared weights of the GEMM that
      // produces y (one fp32 multiply per output element less in the epilogue of every residual layer)
      Buf w1 = prep_weight(base + "conv_res1.weight", Cc, Cc, 9, fold_res ? sum_coeff(u.res_balance) : 1.0);
      const bool has_clip = u.clip_act >= 0.0;
      if (s.heads == 0) {
        ConvArgs a;
        a.res = res;
        a.res_mode = res_mode;
        a.res_rnorm = res_rnorm;
        a.res_t = u.res_balance;
        a.has_clip = has_clip;
        a.clip = u.clip_act;
        a.outs = alloc_outs(out, want(i));
        a.out_rnorm = out.rnorm;
        a.res_folded = fold_res;
        conv(y0, w1, Bc, R, Cc, Cc, 9, a);
      } else {
        Buf xr = a16(Bc, R, Cc);
        temps.push_back(xr);
        {
          ConvArgs a;
          a.res = res;
          a.res_mode = res_mode;
          a.res_rnorm = res_rnorm;
          a.res_t = u.res_balance;
          a.outs = {{xr, VB_OUT_RAW, 1.0}};
          a.res_folded = fold_res;
          conv(y0, w1, Bc, R, Cc, Cc, 9, a);
        }
        const int S = R * R, D = s.head_dim, h = s.heads;
        const int nseg = s.xattn ? feat_seg : 0;
        const int real_seg = zero_feature_keys ? 0 : nseg;
        const int sk = S * (1 + real_seg);
        // D = 32 (the SR UNet): rows zero-padded to 64 elements so that the tcgen05 attention kernel (64-wide operand rows)
        // serves them; the buffers are private to this layer and zeroed once — the GEMM epilogue only ever writes the lower
        // 32 elements.
        const int ld = (D == 32 && S % 256 == 0 && sk % 128 == 0) ? 64 : 0;
        Buf q, k, v;
        if (ld) {
          q = buf(static_cast<long long>(Bc) * h * S * ld, 2, true);
          k = buf(static_cast<long long>(Bc) * h * sk * ld, 2, true);
          v = buf(static_cast<long long>(Bc) * h * sk * ld, 2, true);
        } else {
          q = act(static_cast<long long>(Bc) * h * S, D);
          k = act(static_cast<long long>(Bc) * h * sk, D);
          v = act(static_cast<long long>(Bc) * h * sk, D);
        }
        Buf wq = prep_weight(base + "attn_qkv.weight", 3 * Cc, Cc, 1, 1.0, 0, 3, D);
        {
          QkvArgs qa;
          qa.D = D;
          qa.parts = 3;
          qa.out[0] = q;
          qa.out[1] = k;
          qa.out[2] = v;
          qa.seq[0] = S;
          qa.seq[1] = qa.seq[2] = sk;
          qa.ld = ld;
          ConvArgs a;
          a.qkv = &qa;
          conv(xr, wq, Bc, R, Cc, 3 * Cc, 1, a);
        }
        if (s.xattn && !zero_feature_keys) {
          if (features == nullptr || next_feature >= features->size()) {
            fail("vb_net_plan: %s: no source-view feature map left", s.name.c_str());
            break;
          }
          const Act& f = (*features)[next_feature++];
          if (f.C != Cc || f.R != R) {
            fail("vb_net_plan: %s: feature map mismatch", s.name.c_str());
            break;
          }
          Buf wkv = prep_weight(base + "x_attn_kv.weight", 2 * Cc, Cc, 1, 1.0, 0, 2, D);
          QkvArgs qa;
          qa.D = D;
          qa.parts = 2;
          qa.out[0] = k;
          qa.out[1] = v;
          qa.seq[0] = qa.seq[1] = sk;
          qa.off[0] = qa.off[1] = S;
          qa.seg_div = feat_seg;
          qa.ld = ld;
          ConvArgs a;
          a.qkv = &qa;
          conv(f.raw, wkv, f.B, R, Cc, 2 * Cc, 1, a);
        }
        Buf y = a16(Bc, R, Cc);
        if (ld) {
          temps.push_back(y);            // (padded q/k/v are private: never recycled through the pool)
        } else {
          temps.push_back(q);
          temps.push_back(k);
          temps.push_back(v);
          temps.push_back(y);
        }
        // unconditional model: x_attn_kv(0) == 0 -> the S*nseg zero keys are accounted for analytically
        attention(q, k, v, y, Bc, h, S, sk, D, zero_feature_keys ? S * nseg : 0, ld);
        Buf wp = prep_weight(base + "attn_proj.weight", Cc, Cc, 1, fold_res ? sum_coeff(u.attn_balance) : 1.0);
        ConvArgs a;
        a.res = xr;
        a.res_mode = VB_RES_PLAIN;
        a.res_t = u.attn_balance;
        a.has_clip = has_clip;
        a.clip = u.clip_act;
        a.outs = alloc_outs(out, want(i));
        a.out_rnorm = out.rnorm;
        a.res_folded = fold_res;
        conv(y, wp, Bc, R, Cc, Cc, 1, a);
      }
      if (out.is_feature && feats_out) feats_out->push_back(out);
      if (s.enc_group) skips.push_back(out);
      // recycle: this block's temporaries, the consumed skip, and the previous block's output unless it lives on as a skip
      // connection (encoder outputs) or as a source-view feature
      release(temps);
      if (popped) release(popped_skip.tensors(popped_skip.is_feature));
      if (!cur.is_skip) release(cur.tensors(cur.is_feature));
      cur = out;
    }
    Buf raw;
    if (ok() && u.out_channels > 0) {
      Buf wo = prep_weight(prefix + "out_conv.weight", u.out_channels, cur.C, 9, 1.0, 16, 0, 0, -1, 1.0, 1.0, false, nullptr, prefix + "out_gain");
      raw = buf(static_cast<long long>(Bc) * cur.R * cur.R * 16, 4);
      ConvArgs a;
      a.cout_pad = 16;
      a.out_f32 = raw;
      conv(cur.raw, wo, Bc, cur.R, cur.C, u.out_channels, 9, a);
    }
    return raw;
  }

  // ------------------------------------------------------------------ whole NVPrecond call (engine.Plan._build)
  int enc_ops = 0;
  vb_io_desc io;

  void build() {
    const int R = net.img_resolution;
    const double sd = net.sigma_data;
    const long long img = 3ll * R * R;
    in_x = buf(Bx * img, 4, true);
    if (net.has_encoder) in_src = buf(Bx * img, 4, true);
    in_sigma = buf(Bx, 4);
    fill(in_sigma, 1.0f, Bx);
    const int ldim_enc = net.has_encoder ? net.encoder.label_dim : 0;
    const int ldim_unet = net.unet.label_dim;
    geom_dim = std::max(std::max(ldim_enc, ldim_unet / (net.dual_source ? 2 : 1)), 1);
    in_geom = buf(static_cast<long long>(Bx) * geom_dim, 4, true);
    if (net.super_res) {
      in_cond = buf(B * img, 4, true);
      in_noise = buf(B * img, 4, true);
    }
    out_d = buf(B * img, 4, true);
    const double geom_scale = net.uncond ? 0.0 : 1.0;

    std::vector<Act> features;
    int feat_seg = 1;
    if (net.has_encoder) {
      std::vector<Spec> enc, dec;
      check(unet_layout(net.encoder, enc, dec));
      enc.insert(enc.end(), dec.begin(), dec.end());
      Buf src16 = buf(static_cast<long long>(Bx) * R * R * 64, 2);
      vb_precond_in_desc d;
      memset(&d, 0, sizeof(d));
      d.x = static_cast<const float*>(in_src.ptr);
      d.out = src16.ptr;
      d.B = Bx;
      d.R = R;
      d.cpad = 64;
      d.sigma_n = 1;
      d.sigma_stride = 0;
      d.im2col = 1;
      d.img_stride = img;
      d.sigma_data = static_cast<float>(sd);
      d.noisy_sr = 0.0f;
      emit(5, &d);
      Embedded e = embed(net.encoder, "encoder.", enc, Bx, in_sigma, 1, ldim_enc ? in_geom : Buf(), Bx, ldim_enc, net.no_time_enc ? 0.0 : 1.0,
                         geom_scale);
      run_unet(net.encoder, "encoder.", enc, src16, Bx, e, nullptr, 1, false, true, &features);
      feat_seg = net.dual_source ? 2 : 1;
      if (!dry)
        for (const Act& f : features) plan->features.push_back({f.raw.ptr, f.B, f.R, f.C});
    }
    // ops [0, enc_ops) are the source-view encoder; its outputs are what the reference's return_features / inject_features hand
    // around (training/models.py:664-672, snapshot :612-626)
    enc_ops = num_ops();

    std::vector<Spec> enc, dec;
    check(unet_layout(net.unet, enc, dec));
    enc.insert(enc.end(), dec.begin(), dec.end());
    Buf x16 = buf(static_cast<long long>(B) * R * R * 64, 2);
    const int step = net.dual_source ? 2 : 1;
    {
      vb_precond_in_desc d;
      memset(&d, 0, sizeof(d));
      d.x = static_cast<const float*>(in_x.ptr);
      d.cond = static_cast<const float*>(in_cond.ptr);
      d.noise = static_cast<const float*>(in_noise.ptr);
      d.sigma = static_cast<const float*>(in_sigma.ptr);
      d.out = x16.ptr;
      d.B = B;
      d.R = R;
      d.cpad = 64;
      d.sigma_n = B;
      d.sigma_stride = step;
      d.im2col = 1;
      d.img_stride = img * step;
      d.sigma_data = static_cast<float>(sd);
      d.noisy_sr = static_cast<float>(net.noisy_sr);
      emit(5, &d);
    }
    Embedded e = embed(net.unet, "unet.", enc, B, in_sigma, step, ldim_unet ? in_geom : Buf(), B, ldim_unet, 1.0, geom_scale);
    Buf raw = run_unet(net.unet, "unet.", enc, x16, B, e, &features, feat_seg, !net.has_encoder, false, nullptr);
    if (ok() && !raw) fail("vb_net_plan: the denoising UNet has no out_conv (out_channels == 0)");
    {
      vb_precond_out_desc d;
      memset(&d, 0, sizeof(d));
      d.x = static_cast<const float*>(in_x.ptr);
      d.f = static_cast<const float*>(raw.ptr);
      d.sigma = static_cast<const float*>(in_sigma.ptr);
      d.d_out = static_cast<float*>(out_d.ptr);
      d.B = B;
      d.R = R;
      d.ldf = 16;
      d.sigma_n = B;
      d.sigma_stride = step;
      d.img_stride = img * step;
      d.sigma_data = static_cast<float>(sd);
      emit(6, &d);
    }
    memset(&io, 0, sizeof(io));
    io.in_x = static_cast<float*>(in_x.ptr);
    io.in_src = static_cast<float*>(in_src.ptr);
    io.in_sigma = static_cast<float*>(in_sigma.ptr);
    io.in_geom = static_cast<float*>(in_geom.ptr);
    io.in_cond = static_cast<float*>(in_cond.ptr);
    io.in_noise = static_cast<float*>(in_noise.ptr);
    io.out_d = static_cast<float*>(out_d.ptr);
    io.n_x = Bx;
    io.n_out = B;
    io.img_elems = img;
    io.geom_dim = geom_dim;
    io.workspace_bytes = owned_bytes;
    if (!ok()) return;
    if (dry) {
      trace.add("enc_ops %d\n", enc_ops);
      format_desc(trace, 7, &io);
    } else {
      check(vb_plan_bind_io(plan, &io));
      plan->enc_ops = enc_ops;
    }
  }
};

int check_net(const vb_net_desc* net, const vb_param* params, int32_t n_params, int32_t batch) {
  VB_REQUIRE(net != nullptr && params != nullptr && n_params > 0, "vb_net_plan: net and params are required");
  VB_REQUIRE(batch > 0, "vb_net_plan: batch must be positive");
  VB_REQUIRE(net->img_resolution >= 4 && (net->img_resolution & (net->img_resolution - 1)) == 0 &&
                 net->unet.img_resolution == net->img_resolution && (!net->has_encoder || net->encoder.img_resolution == net->img_resolution),
             "vb_net_plan: img_resolution must be a power of two >= 4, the same for the net and its UNets");
  VB_REQUIRE(net->unet.out_channels == 3, "vb_net_plan: the denoising UNet must have 3 output channels");
  VB_REQUIRE(!net->has_encoder || net->encoder.out_channels == 0, "vb_net_plan: the source-view encoder is a UNetEncoder (out_channels = 0)");
  VB_REQUIRE((net->uncond != 0) == (net->has_encoder == 0), "vb_net_plan: uncond nets have no source-view encoder, all others have one");
  VB_REQUIRE(net->unet.in_channels == (net->super_res ? 7 : 4), "vb_net_plan: in_channels must be 4 (7 for super_res nets)");
  VB_REQUIRE(!net->dual_source || net->unet.label_dim == 2 * net->encoder.label_dim, "vb_net_plan: dual-source nets expect target_label_dim == 2 * source_label_dim");
  for (int i = 0; i < n_params; ++i)
    VB_REQUIRE(params[i].name != nullptr && params[i].data != nullptr && params[i].ndim >= 0 && params[i].ndim <= 4, "vb_net_plan: bad entry %d of the parameter table", i);
  return VB_OK;
}

}  // namespace

extern "C" int vb_net_plan_create(const vb_net_desc* net, const vb_param* params, int32_t n_params, int32_t batch, void* stream, vb_plan** out) {
  VB_REQUIRE(out != nullptr, "vb_net_plan_create: null out");
  *out = nullptr;
  int rc = check_net(net, params, n_params, batch);
  if (rc != VB_OK) return rc;
  rc = vb_device_check();
  if (rc != VB_OK) return rc;
  vb_plan* plan = nullptr;
  rc = vb_plan_create(&plan);
  if (rc != VB_OK) return rc;
  Recorder r(*net, params, n_params, batch, false, static_cast<cudaStream_t>(stream));
  r.plan = plan;
  r.build();
  if (r.ok()) r.check_cuda(cudaStreamSynchronize(r.stream), "cudaStreamSynchronize");      // weight preparation read the caller's tensors
  if (!r.ok()) {
    vb_plan_destroy(plan);
    return r.rc;
  }
  *out = plan;
  return VB_OK;
}

// Refresh every prepared weight of a library-recorded plan from another parameter table of the same architecture.
extern "C" int vb_net_plan_set_weights(vb_plan* p, const vb_param* params, int32_t n_params, void* stream) {
  VB_REQUIRE(p != nullptr && params != nullptr && n_params > 0, "vb_net_plan_set_weights: plan and params are required");
  VB_REQUIRE(!p->weights.empty(), "vb_net_plan_set_weights: the plan was not recorded by vb_net_plan_create");
  std::unordered_map<std::string, const vb_param*> index;
  for (int i = 0; i < n_params; ++i) {
    VB_REQUIRE(params[i].name != nullptr && params[i].data != nullptr && params[i].ndim >= 0 && params[i].ndim <= 4,
               "vb_net_plan_set_weights: bad entry %d of the parameter table", i);
    index[params[i].name] = &params[i];
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto lookup = [&](const std::string& name, long long numel) -> const vb_param* {
    auto it = index.find(name);
    if (it == index.end()) {
      vb::set_error("vb_net_plan_set_weights: parameter '%s' is missing", name.c_str());
      return nullptr;
    }
    long long n = 1;
    for (int i = 0; i < it->second->ndim; ++i) n *= it->second->shape[i];
    if (n != numel || (it->second->dtype != VB_F32 && it->second->dtype != VB_F16)) {
      vb::set_error("vb_net_plan_set_weights: '%s' must hold %lld fp32/fp16 values", name.c_str(), numel);
      return nullptr;
    }
    return it->second;
  };
  // validate the whole table before the first launch: a refresh is all or nothing
  for (const vb_plan::WeightSlot& w : p->weights) {
    if (lookup(w.name, w.numel) == nullptr) return VB_ERR_INVALID;
    if (!w.gain_param.empty() && lookup(w.gain_param, 1) == nullptr) return VB_ERR_INVALID;
  }
  for (const vb_plan::WeightSlot& w : p->weights) {
    const vb_param* src = lookup(w.name, w.numel);
    if (w.copy) {
      to_f32_kernel<<<static_cast<unsigned>((w.numel + 255) / 256), 256, 0, s>>>(src->data, src->dtype, static_cast<float*>(w.d.dst), w.numel);
      VB_CHECK_CUDA(cudaGetLastError());
      continue;
    }
    vb_weight_prep_desc d = w.d;
    d.src = src->data;
    d.src_dtype = src->dtype;
    double gain = w.gain_const;
    if (!w.gain_param.empty()) {
      const vb_param* g = lookup(w.gain_param, 1);
      unsigned char raw[4] = {0, 0, 0, 0};
      VB_CHECK_CUDA(cudaMemcpyAsync(raw, g->data, g->dtype == VB_F16 ? 2 : 4, cudaMemcpyDeviceToHost, s));
      VB_CHECK_CUDA(cudaStreamSynchronize(s));
      if (g->dtype == VB_F16) {
        __half h;
        memcpy(&h, raw, 2);
        gain *= static_cast<double>(__half2float(h));
      } else {
        float f;
        memcpy(&f, raw, 4);
        gain *= static_cast<double>(f);
      }
    }
    d.gain = static_cast<float>(gain);
    const int rc = vb_weight_prep(&d, stream);
    if (rc != VB_OK) return rc;
  }
  return VB_OK;
}

extern "C" int vb_plan_get_io(const vb_plan* p, vb_io_desc* out, int32_t* enc_ops) {
  VB_REQUIRE(p != nullptr && p->io_bound && out != nullptr, "vb_plan_get_io: the plan has no bound I/O buffers");
  *out = p->io;
  if (enc_ops != nullptr) *enc_ops = p->enc_ops;
  return VB_OK;
}

extern "C" int vb_plan_num_features(const vb_plan* p) { return p ? static_cast<int>(p->features.size()) : 0; }

extern "C" int vb_plan_get_feature(const vb_plan* p, int32_t i, void** ptr, int32_t* B, int32_t* R, int32_t* C) {
  VB_REQUIRE(p != nullptr && ptr != nullptr && B != nullptr && R != nullptr && C != nullptr, "vb_plan_get_feature: null argument");
  VB_REQUIRE(i >= 0 && i < static_cast<int>(p->features.size()), "vb_plan_get_feature: the plan has %d feature maps", static_cast<int>(p->features.size()));
  const vb_plan::Feature& f = p->features[i];
  *ptr = f.ptr;
  *B = f.B;
  *R = f.R;
  *C = f.C;
  return VB_OK;
}

extern "C" int64_t vb_net_plan_trace(const vb_net_desc* net, const vb_param* params, int32_t n_params, int32_t batch, char* buf, int64_t cap) {
  int rc = check_net(net, params, n_params, batch);
  if (rc != VB_OK) return rc;
  Recorder r(*net, params, n_params, batch, true, nullptr);
  r.build();
  if (!r.ok()) return r.rc;
  const std::string& s = r.trace.s;
  if (buf != nullptr && cap > 0) {
    const size_t n = std::min(s.size(), static_cast<size_t>(cap - 1));
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return static_cast<int64_t>(s.size());
}

extern "C" int64_t vb_trace_desc(int32_t kind, const void* desc, char* buf, int64_t cap) {
  if (desc == nullptr || kind < 0 || kind > 7) {
    vb::set_error("vb_trace_desc: bad argument");
    return VB_ERR_INVALID;
  }
  Text t;
  format_desc(t, kind, desc);
  if (buf != nullptr && cap > 0) {
    const size_t n = std::min(t.s.size(), static_cast<size_t>(cap - 1));
    memcpy(buf, t.s.data(), n);
    buf[n] = 0;
  }
  return static_cast<int64_t>(t.s.size());
}
