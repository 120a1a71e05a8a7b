"""Plans recorded INSIDE the library (include/vivid_b200.h: vb_net_plan_create — SURVEY.md 8(b) `vb_plan_create(net_desc)` +
`vb_plan_set_weights(names, ptrs)`), and the op-for-op comparison between that recorder and `engine.Plan`.

  net_desc(net) / param_table(net)   the C description of an NVPrecond (constructor arguments; parameters under their
                                     reference state_dict names) — what a non-Python host would fill from a checkpoint
  LibPlan(net, batch, device)        a plan recorded by libvividb200.so itself (owns its device buffers); runs through
                                     vb_denoise / vb_sample
  trace_library(net, batch)          dry run of the library's recorder (no device): one text line per allocation, weight
                                     preparation and recorded op, pointers canonicalised
  trace_engine(net, batch)           the same text from engine.Plan's walk (its library calls are intercepted, buffers are host
                                     tensors); tests/test_netplan.py asserts the two are identical for every preset
"""
import bisect
import ctypes as C

import torch

from . import _lib as L
from . import engine

_ADDR_SHIFT = 36
_PARAM_BASE = 0x4000


def _unet_desc(unet, block, xattn, out_channels, in_channels):
    d = L.UNetDesc(img_resolution=unet.img_resolution, in_channels=in_channels, out_channels=out_channels,
                   model_channels=unet.model_channels, num_levels=len(unet.channel_mult), num_blocks=unet.num_blocks,
                   num_attn_res=len(unet.attn_resolutions), extra_attn=-1 if unet.extra_attn is None else int(unet.extra_attn),
                   channels_per_head=unet.channels_per_head, xattn=int(xattn), label_dim=unet.label_dim, cnoise=unet.cnoise,
                   cemb=unet.cemb, label_balance=float(unet.label_balance), concat_balance=float(unet.concat_balance),
                   res_balance=float(block.res_balance), attn_balance=float(block.attn_balance),
                   clip_act=-1.0 if block.clip_act is None else float(block.clip_act))
    for i, m in enumerate(unet.channel_mult):
        d.channel_mult[i] = int(m)
    for i, r in enumerate(unet.attn_resolutions):
        d.attn_resolutions[i] = int(r)
    return d


def net_desc(net):
    """vb_net_desc of a vivid_b200.NVPrecond (the reference's constructor arguments, training/models.py:589-627)."""
    def first_block(u):
        return next(m for s, m in ((s, (u.enc if s.group == "enc" else u.dec)[s.name]) for s in u.enc_specs + u.dec_specs) if s.kind == "block")
    d = L.NetDesc(has_encoder=int(net.encoder is not None), img_resolution=net.img_resolution, uncond=int(bool(net.uncond)),
                  super_res=int(bool(net.super_res)), dual_source=int(bool(net.dual)), no_time_enc=int(bool(net.no_time_enc)),
                  sigma_data=float(net.sigma_data), noisy_sr=float(net.noisy_sr if net.noisy_sr is not None else 0.0))
    d.unet = _unet_desc(net.unet, first_block(net.unet), True, 3, net.img_channels + 1 + (net.img_channels if net.super_res else 0))
    if net.encoder is not None:
        d.encoder = _unet_desc(net.encoder, first_block(net.encoder), False, 0, net.img_channels + 1)
    return d


_DT = {torch.float32: L.VB_F32, torch.float16: L.VB_F16}


def param_table(net):
    """(array of vb_param, keep-alive list): every parameter and buffer under its state_dict name, in state_dict order."""
    items = [(k, v) for k, v in net.state_dict().items()]
    arr = (L.Param * len(items))()
    keep = []
    for i, (k, v) in enumerate(items):
        if v.dtype not in _DT:
            raise TypeError(f"{k}: parameters must be fp32 or fp16, got {v.dtype}")
        v = v.detach()
        if not v.is_contiguous():
            v = v.contiguous()
        name = k.encode()
        keep += [v, name]
        arr[i].name = name
        arr[i].data = v.data_ptr()
        arr[i].dtype = _DT[v.dtype]
        arr[i].ndim = v.ndim
        for j, s in enumerate(v.shape):
            arr[i].shape[j] = s
    return arr, keep


class _DeviceMemory:
    """A library-owned device buffer as seen by torch (CUDA array interface): torch.as_tensor(...) wraps it without a copy."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def _view(ptr, shape, dtype, device):
    if not ptr:
        return None
    if dtype == torch.float32:
        return torch.as_tensor(_DeviceMemory(ptr, shape, "<f4"), device=device)
    return torch.as_tensor(_DeviceMemory(ptr, shape, "<i2"), device=device).view(dtype)      # fp16 / bf16 operands


class LibPlan:
    """A denoiser plan recorded by the library from (net description, parameter table): vb_net_plan_create.  Offers what
    NVPrecond.forward / edm_sampler use of engine.Plan (I/O tensors — views of the plan's own device buffers —, run, feature
    maps), so `VB_LIB_PLAN=1` makes the whole Python surface run on library-recorded plans."""

    def __init__(self, net, batch, device):
        self.lib = L.lib()
        self.handle = C.c_void_p()
        self.net, self.B, self.device = net, int(batch), torch.device(device)
        desc = net_desc(net)
        params, keep = param_table(net)
        stream = torch.cuda.current_stream(device).cuda_stream
        with torch.cuda.device(device):
            L.check(self.lib.vb_net_plan_create(C.byref(desc), params, len(params), int(batch), stream, C.byref(self.handle)),
                    "vb_net_plan_create")
        del keep
        self.io = L.IoDesc()
        enc_ops = C.c_int32()
        L.check(self.lib.vb_plan_get_io(self.handle, C.byref(self.io), C.byref(enc_ops)), "vb_plan_get_io")
        self.enc_ops = enc_ops.value
        self.num_ops = self.lib.vb_plan_num_ops(self.handle)
        self.launches = int(self.lib.vb_plan_query(self.handle, 1))
        self.padded_flops = self.lib.vb_plan_query(self.handle, 0)
        self.owned_bytes = int(self.io.workspace_bytes)
        self.weight_versions = None
        io, R, f32 = self.io, net.img_resolution, torch.float32
        self.in_x = _view(io.in_x, (io.n_x, 3, R, R), f32, device)
        self.in_src = _view(io.in_src, (io.n_x, 3, R, R), f32, device)
        self.in_sigma = _view(io.in_sigma, (io.n_x,), f32, device)
        self.in_geom = _view(io.in_geom, (io.n_x, io.geom_dim), f32, device)
        self.in_cond = _view(io.in_cond, (io.n_out, 3, R, R), f32, device)
        self.in_noise = _view(io.in_noise, (io.n_out, 3, R, R), f32, device)
        self.out_d = _view(io.out_d, (io.n_out, 3, R, R), f32, device)
        self._features = []
        op_dtype = L.operand_torch_dtype()
        for i in range(self.lib.vb_plan_num_features(self.handle)):
            ptr, fb, fr, fc = C.c_void_p(), C.c_int32(), C.c_int32(), C.c_int32()
            L.check(self.lib.vb_plan_get_feature(self.handle, i, C.byref(ptr), C.byref(fb), C.byref(fr), C.byref(fc)), "vb_plan_get_feature")
            self._features.append(_view(ptr.value, (fb.value, fr.value, fr.value, fc.value), op_dtype, device))

    def set_weights(self, net):
        """Refresh the plan's prepared weights from `net` (same architecture; e.g. after load_state_dict): vb_net_plan_set_weights."""
        params, keep = param_table(net)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            L.check(self.lib.vb_net_plan_set_weights(self.handle, params, len(params), stream), "vb_net_plan_set_weights")
            torch.cuda.current_stream(self.device).synchronize()       # the table's tensors may go away with `keep`
        del keep

    def run(self, graph=True, section="all"):
        """engine.Plan.run: everything, the source-view encoder alone ('enc') or the denoising UNet alone ('unet')."""
        first, last = {"all": (0, -1), "enc": (0, self.enc_ops), "unet": (self.enc_ops, -1)}[section]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        fn = self.lib.vb_plan_launch_graph_range if graph else self.lib.vb_plan_run
        L.check(fn(self.handle, first, last, stream), "vb_plan_run")

    def feature_views(self):
        """The encoder's cross-attention feature maps as logical NCHW views of the plan's 16-bit NHWC buffers."""
        return [f.permute(0, 3, 1, 2) for f in self._features]

    def __del__(self):
        try:
            if self.handle:
                self.lib.vb_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------ op-for-op comparison
def trace_library(net, batch):
    """Text trace of the library recorder's dry run for `net` (host tensors are fine: no device is touched)."""
    lib = L.lib()
    desc = net_desc(net)
    params, keep = param_table(net)
    n = lib.vb_net_plan_trace(C.byref(desc), params, len(params), int(batch), None, 0)
    if n < 0:
        L.check(int(n), "vb_net_plan_trace")
    buf = C.create_string_buffer(n + 1)
    lib.vb_net_plan_trace(C.byref(desc), params, len(params), int(batch), buf, n + 1)
    del keep
    return buf.value.decode()


class _Addresses:
    """Real host pointers -> the canonical addresses of a dry run (buffer k at (k+1) << 36, parameter i at (0x4000+i) << 36)."""

    def __init__(self):
        self.starts, self.spans = [], []

    def add(self, ptr, nbytes, canon):
        i = bisect.bisect_left(self.starts, ptr)
        self.starts.insert(i, ptr)
        self.spans.insert(i, (ptr, max(nbytes, 1), canon))

    def __call__(self, ptr):
        if not ptr:
            return None
        i = bisect.bisect_right(self.starts, ptr) - 1
        if i >= 0:
            start, nbytes, canon = self.spans[i]
            if start <= ptr < start + nbytes:
                return canon + (ptr - start)
        raise AssertionError(f"pointer {ptr:#x} belongs to no buffer or parameter of the plan")


def _fmt_ptr(key, a):
    if not a:
        return f" {key}=-"
    hi, off = a >> _ADDR_SHIFT, a & ((1 << _ADDR_SHIFT) - 1)
    return f" {key}=p{hi - _PARAM_BASE}+{off}" if hi >= _PARAM_BASE else f" {key}=b{hi - 1}+{off}"


class _TraceLib:
    """Stands in for the ctypes library while engine.Plan walks a net on the host: every recording call becomes a text line
    (formatted by the library's own vb_trace_desc after the pointers have been canonicalised)."""
    KIND = {"vb_weight_prep": 0, "vb_plan_add_conv": 1, "vb_plan_add_attn": 2, "vb_plan_add_eltwise": 3, "vb_plan_add_embed": 4,
            "vb_plan_add_precond_in": 5, "vb_plan_add_precond_out": 6, "vb_plan_bind_io": 7}

    def __init__(self, plan):
        self.real = L.lib()
        self.plan = plan
        self.ops = 0

    def line(self, kind, desc):
        d = type(desc).from_buffer_copy(desc)
        for name, ctype in d._fields_:
            if ctype is C.c_void_p:
                setattr(d, name, self.plan.addr(getattr(d, name)))
            elif isinstance(ctype, type) and issubclass(ctype, C.Array) and ctype._type_ is C.c_void_p:
                arr = getattr(d, name)
                for j in range(len(arr)):
                    arr[j] = self.plan.addr(arr[j])
        buf = C.create_string_buffer(4096)
        n = self.real.vb_trace_desc(kind, C.byref(d), buf, len(buf))
        assert 0 < n < len(buf)
        self.plan.lines.append(buf.value.decode())

    def __getattr__(self, name):
        if name in self.KIND:
            kind = self.KIND[name]

            def record(*args):
                desc = args[1 if name.startswith("vb_plan") else 0]._obj
                if kind == 7:
                    self.plan.lines.append(f"enc_ops {self.plan.enc_ops}\n")
                elif kind >= 1:
                    self.ops += 1
                self.line(kind, desc)
                return 0
            return record
        if name in ("vb_device_check", "vb_plan_create"):
            return lambda *a: 0
        if name == "vb_plan_destroy":
            return lambda *a: None
        if name == "vb_plan_num_ops":
            return lambda *a: self.ops
        if name == "vb_plan_query":
            return lambda *a: 0.0
        if name in ("vb_operand_dtype", "vb_conv_ksplit_ws_bytes"):
            return getattr(self.real, name)
        raise AttributeError(f"engine.Plan called {name}, which the tracer does not model")


class _TracePlan(engine.Plan):
    def __init__(self, net, B):
        self.lines = []
        self.addr = _Addresses()
        self.n_bufs = 0
        for i, (k, v) in enumerate(net.state_dict().items()):
            self.addr.add(v.data_ptr(), v.numel() * v.element_size(), (_PARAM_BASE + i) << _ADDR_SHIFT)
        self.lib = _TraceLib(self)
        self.op_dtype = L.operand_torch_dtype()
        self.op_code = self.lib.vb_operand_dtype()
        self.net, self.device, self.B = net, torch.device("cpu"), B
        self.dual = net.dual
        self.Bx = 2 * B if net.dual else B
        self.keep, self.owned_bytes, self.pool, self.op_info = [], 0, {}, []
        self.handle, self.stream, self.sm_count = None, None, 148
        self.alg_flops, self.weight_versions, self.ks_ws = 0.0, None, None
        self._build()

    def buf(self, shape, dtype, zero=False):
        t = super().buf(shape, dtype, zero)
        nbytes = t.numel() * t.element_size()
        self.lines.append(f"alloc b{self.n_bufs} {nbytes} z{1 if zero else 0}\n")
        self.addr.add(t.data_ptr(), nbytes, (self.n_bufs + 1) << _ADDR_SHIFT)
        self.n_bufs += 1
        return t

    def to_f32(self, dst, src):
        self.lines.append("to_f32" + _fmt_ptr("dst", self.addr(dst.data_ptr())) + _fmt_ptr("src", self.addr(src.data_ptr())) +
                          f" n={src.numel()} dt={_DT[src.dtype]}\n")

    def fill(self, dst, value):
        self.lines.append("fill" + _fmt_ptr("dst", self.addr(dst.data_ptr())) + f" v={value:.9g} n={dst.numel()}\n")


def trace_engine(net, batch):
    """Text trace of engine.Plan's walk over `net` (host tensors; layout tuning off: the heuristic N tile is what both recorders
    start from)."""
    saved = engine.AUTOTUNE
    engine.AUTOTUNE = False
    try:
        return "".join(_TracePlan(net, int(batch)).lines)
    finally:
        engine.AUTOTUNE = saved
