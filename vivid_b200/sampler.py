"""EDM Heun sampler with autoguidance — drop-in for the reference's `edm_sampler`
(generate_images.py:43-118; vanilla form experiments/code/generate_images.py:41-91).

Same keyword signature, so it can be passed as `sampler_fn=` to generate_images_nvs.  The sigma
schedule is computed with the reference's exact fp32 torch expression; the guidance lerp and both
Heun updates run in one fused CUDA pass each (vb_heun) instead of ~10 eager pointwise launches.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L


def _heun(lib, d_net, d_gnet, x_hat, d_cur, x_next, phase, guidance, t_hat, t_next, x_out=(), sigma_out=(), sigma_next=0.0):
    desc = L.HeunDesc(d_net=d_net.data_ptr(), d_gnet=L.ptr(d_gnet), x_hat=x_hat.data_ptr(), d_cur=d_cur.data_ptr(),
                      x_next=x_next.data_ptr(), n=x_hat.numel(), phase=phase, guidance=float(guidance),
                      t_hat=float(t_hat), t_next=float(t_next), sigma_next=float(sigma_next),
                      sigma_n=sigma_out[0].numel() if sigma_out else 0)
    for k, t in enumerate(x_out):
        desc.x_out[k] = t.data_ptr()
    for k, t in enumerate(sigma_out):
        desc.sigma_out[k] = t.data_ptr()
    L.check(lib.vb_heun(C.byref(desc), torch.cuda.current_stream(x_hat.device).cuda_stream), "vb_heun")


_SIDE = {}


def _side_stream(device):
    key = str(device)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device)
    return _SIDE[key]


def sigma_steps(num_steps, sigma_min, sigma_max, rho, device, dtype=torch.float32):
    """t_i = (smax^(1/rho) + i/(N-1) (smin^(1/rho) - smax^(1/rho)))^rho, t_N = 0  (generate_images.py:68-70)."""
    step_indices = torch.arange(num_steps, dtype=dtype, device=device)
    t = (sigma_max ** (1 / rho) + step_indices / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    return torch.cat([t, torch.zeros_like(t[:1])])


def edm_sampler(net, src, noise, labels=None, gnet=None, conditioning_image=None, num_steps=32, sigma_min=0.002,
                sigma_max=80, rho=7, guidance=1, S_churn=0, S_min=0, S_max=float("inf"), S_noise=1,
                dtype=torch.float32, randn_like=torch.randn_like, _trace=None):
    """`_trace` (testing aid, not part of the reference's signature): a list that receives a copy of the guided
    denoiser output D of every call, in call order."""
    if dtype != torch.float32:
        raise NotImplementedError("the sampler state is fp32 (the reference default); other dtypes are not implemented")
    if noise.device.type != "cuda":
        raise RuntimeError("vivid_b200.edm_sampler runs on CUDA only; there is no CPU fallback")
    with torch.cuda.device(noise.device):        # kernels launch on the tensors' device whatever the caller's current one is
        if _bound_ok(net, gnet, guidance, S_churn):
            return _edm_sampler_bound(net, src, noise, labels, gnet, conditioning_image, num_steps, sigma_min, sigma_max,
                                      rho, guidance, dtype, _trace)
        return _edm_sampler_generic(net, src, noise, labels, gnet, conditioning_image, num_steps, sigma_min, sigma_max, rho,
                                    guidance, S_churn, S_min, S_max, S_noise, dtype, randn_like, _trace)


def _bound_ok(net, gnet, guidance, S_churn):
    """The zero-copy loop applies to this package's own fp16-path nets with vanilla semantics; everything else (wrapped
    callables, hooks, dual-source nets, fp32 validation mode, the stochastic branch) takes the generic loop."""
    from .precond import NVPrecond
    if os.environ.get("VB_BOUND_SAMPLER", "1") == "0" or S_churn > 0:
        return False

    def ok(n):
        return (type(n) is NVPrecond and not n.dual and n.use_fp16 and not n._forward_hooks and not n._forward_pre_hooks)
    if not ok(net):
        return False
    if getattr(net, "no_time_enc", None) and (net.encoder is None or net.super_res):
        return False                                    # the generic loop raises the reference-shaped error
    if guidance != 1:
        if gnet is net or not ok(gnet) or gnet.super_res or gnet.img_resolution != net.img_resolution:
            return False
    return True


class _Bound:
    """One net bound to the inputs that stay constant over a sampler call (source view, pose vector, SR conditioning
    image): they are uploaded into the plan's buffers ONCE, the plan is built and tuned on the caller's stream, and each
    denoiser call is then a graph replay whose x / sigma inputs were written by the previous vb_heun pass."""

    def __init__(self, net, src, labels, cond, n):
        dev = src.device
        self.net, self.p = net, net.plan(n, dev)
        p = self.p
        if p.in_src is not None:
            if src.shape != p.in_src.shape:
                raise ValueError(f"src must be {tuple(p.in_src.shape)}, got {tuple(src.shape)}")
            p.in_src.copy_(src)
        if labels is None:
            p.in_geom.zero_()
        else:
            g = labels.to(torch.float32).reshape(-1, p.in_geom.shape[1])
            p.in_geom.copy_(g.expand(n, -1) if g.shape[0] == 1 else g)
        if net.super_res:
            if cond is None:
                raise AssertionError("super_res model requires conditioning_image")
            p.in_cond.copy_(cond)
        self.section = "all"

    def cache_features(self, sigma0):
        """no_time_enc nets (generate_images.py:52-57): the source-view encoder runs once per sampler call."""
        self.p.in_sigma.fill_(sigma0)
        self.p.run(graph=self.net.use_graph, section="enc")
        self.section = "unet"

    def launch(self):
        if self.net.super_res:
            # the reference draws this from the GLOBAL generator on every call; normal_() on the plan's
            # buffer consumes the generator exactly as torch.randn_like(conditioning_image) does
            self.p.in_noise.normal_()
        self.p.run(graph=self.net.use_graph, section=self.section)
        return self.p.out_d


def _sample_c(lib, bn, bg, noise, t_steps, guidance, side, cur):
    """vb_sample over the bound plans.  SR nets draw their per-call conditioning noise through a callback that runs the
    reference's own draw (normal_() on the plan's buffer == torch.randn_like from the global generator)."""
    p = bn.p
    ws = torch.empty(lib.vb_sample_workspace_bytes(p.handle) // 4, dtype=torch.float32, device=noise.device)
    out = torch.empty_like(noise)
    failure = []

    def draw(user, dst, n, stream):
        try:
            assert dst == p.in_noise.data_ptr() and n == p.in_noise.numel()
            p.in_noise.normal_()
            return 0
        except BaseException as e:      # never let an exception cross the C frame
            failure.append(e)
            return -1
    d = L.SampleDesc(net=p.handle, gnet=bg.p.handle if bg is not None else None, noise=noise.data_ptr(),
                     t_steps=(C.c_float * len(t_steps))(*t_steps), workspace=ws.data_ptr(), x_out=out.data_ptr(),
                     side_stream=side.cuda_stream if side is not None else None, num_steps=len(t_steps) - 1,
                     net_first_op=p.enc_ops if bn.section == "unet" else 0, guidance=float(guidance))
    if bn.net.super_res:
        d.sr_noise = L.NOISE_FN(draw)
    rc = lib.vb_sample(C.byref(d), cur.cuda_stream)
    if failure:
        raise failure[0]
    L.check(rc, "vb_sample")
    return out


def _edm_sampler_bound(net, src, noise, labels, gnet, cond, num_steps, sigma_min, sigma_max, rho, guidance, dtype, trace):
    lib = L.lib()
    dev = noise.device
    n = noise.shape[0]
    t_dev = sigma_steps(num_steps, sigma_min, sigma_max, rho, dev, dtype)
    t_steps = t_dev.tolist()                       # one host sync per sampler call
    guided = guidance != 1
    bn = _Bound(net, src, labels, cond, n)
    bg = _Bound(gnet, src, None, None, n) if guided else None       # gnet(src, x, t): no pose, no conditioning image
    if getattr(net, "no_time_enc", None):
        bn.cache_features(t_steps[0])
    bounds = [bn] + ([bg] if guided else [])
    x_in = [b.p.in_x for b in bounds]
    sig_in = [b.p.in_sigma for b in bounds]
    side = _side_stream(dev) if guided and os.environ.get("VB_DUAL_STREAM", "1") != "0" else None
    cur = torch.cuda.current_stream(dev)
    # (a degenerate schedule — num_steps = 1 is 0/0 in the reference's formula, generate_images.py:69 — takes the Python loop and
    #  yields the reference's NaNs instead of vb_sample's argument error)
    if (trace is None and os.environ.get("VB_C_SAMPLER", "1") != "0" and all(b.net.use_graph for b in bounds)
            and all(t > 0 for t in t_steps[:-1])):
        # the whole loop in the library (vb_sample: the same graph replays and vb_heun passes, enqueued from C) — what a
        # non-Python host would call; the Python loop below is the same sequence and stays for tracing / eager replay
        return _sample_c(lib, bn, bg, noise.to(dtype).contiguous(), t_steps, guidance, side, cur)
    x_hat = (noise.to(dtype) * t_dev[0]).contiguous()
    x_next, d_cur = torch.empty_like(x_hat), torch.empty_like(x_hat)
    for b in bounds:
        b.p.in_x.copy_(x_hat)
        b.p.in_sigma.fill_(t_steps[0])

    def denoise():
        if side is None:
            dn = bn.launch()
            return dn, (bg.launch() if guided else None)
        side.wait_stream(cur)                      # both nets' inputs are in place (written by the previous vb_heun)
        dn = bn.launch()
        with torch.cuda.stream(side):
            dg = bg.launch()
        cur.wait_stream(side)
        return dn, dg

    def log(dn, dg):
        if trace is not None:
            trace.append(dn.clone() if dg is None else torch.lerp(dg, dn, guidance))

    for i in range(num_steps):
        t_hat, t_nxt = t_steps[i], t_steps[i + 1]
        last = i == num_steps - 1
        dn, dg = denoise()
        log(dn, dg)
        # Euler step; x_next and the next call's noise level go straight into the plans' input buffers
        _heun(lib, dn, dg, x_hat, d_cur, x_next, 0, guidance, t_hat, t_nxt, x_out=() if last else x_in,
              sigma_out=() if last else sig_in, sigma_next=t_nxt)
        if not last:
            dn, dg = denoise()
            log(dn, dg)
            _heun(lib, dn, dg, x_hat, d_cur, x_next, 1, guidance, t_hat, t_nxt, x_out=x_in, sigma_out=sig_in,
                  sigma_next=t_nxt)                   # 2nd-order correction
        x_hat, x_next = x_next, x_hat
    return x_hat


def _edm_sampler_generic(net, src, noise, labels, gnet, conditioning_image, num_steps, sigma_min, sigma_max, rho, guidance,
                         S_churn, S_min, S_max, S_noise, dtype, randn_like, trace):
    lib = L.lib()
    dual = bool(getattr(net, "dual", False))
    t_dev = sigma_steps(num_steps, sigma_min, sigma_max, rho, noise.device, dtype)
    t_steps = t_dev.tolist()                       # one host sync per sampler call

    # a net whose source-view encoder ignores the noise level runs it once per batch (generate_images.py:52-57)
    features = None
    if getattr(net, "no_time_enc", None):
        features = net(src, torch.zeros_like(src), torch.ones(src.shape[0], dtype=dtype, device=noise.device), labels,
                       conditioning_image, return_features=True)

    # The guiding net's call is independent of the main net's (generate_images.py:57-62): replay the two plans on two
    # streams so that the tail, set-up and single-wave layers of one overlap the other (VB_DUAL_STREAM=0: one stream).
    side = None
    if guidance != 1 and gnet is not net and os.environ.get("VB_DUAL_STREAM", "1") != "0":
        side = _side_stream(noise.device)
        for m in (net, gnet):                      # build and tune both plans on the caller's stream, not concurrently
            if hasattr(m, "plan") and getattr(m, "use_fp16", False):
                m.plan(noise.shape[0] // (2 if getattr(m, "dual", False) else 1), noise.device)
        torch.cuda.current_stream(noise.device).synchronize()

    def denoise(x, t):
        tt = torch.full((x.shape[0],), t, dtype=dtype, device=x.device)
        if side is None:
            dn = net(src, x, tt, labels, conditioning_image, inject_features=features)
            dg = gnet(src, x, tt) if guidance != 1 else None
            return dn, dg
        cur = torch.cuda.current_stream(x.device)
        side.wait_stream(cur)                     # the inputs are ready; recorded BEFORE net's launch, so gnet does not wait for it
        dn = net(src, x, tt, labels, conditioning_image, inject_features=features)
        with torch.cuda.stream(side):             # host order net -> gnet as in the reference (global-RNG draws of SR nets, F7)
            dg = gnet(src, x, tt)
        cur.wait_stream(side)
        dg.record_stream(cur)
        for t_in in (src, x, tt):                 # read on the side stream: keep the allocator from recycling them early
            t_in.record_stream(side)
        return dn, dg

    def widen(xh):                                 # dual-source: every target appears twice (generate_images.py:96-98)
        return xh.repeat_interleave(2, dim=0) if dual else xh

    x_next = (noise.to(dtype) * t_dev[0])
    if dual:
        x_next = x_next[::2]
    x_next = x_next.contiguous()
    d_cur = torch.empty_like(x_next)
    for i in range(num_steps):
        t_cur, t_nxt = t_steps[i], t_steps[i + 1]
        x_cur = x_next
        if S_churn > 0 and S_min <= t_cur <= S_max:
            gamma = min(S_churn / num_steps, np.sqrt(2) - 1)
            t_hat = float(np.float32(t_cur) + np.float32(gamma) * np.float32(t_cur))
            # (dual-source: the reference draws for the 2B interleaved state and keeps the even rows, :80,:92)
            eps = randn_like(widen(x_cur))[::2] if dual else randn_like(x_cur)
            coef = np.sqrt(np.float32(t_hat) ** 2 - np.float32(t_cur) ** 2) * np.float32(S_noise)     # fp32 like the reference
            x_hat = x_cur + float(coef) * eps
        else:
            t_hat, x_hat = t_cur, x_cur
        dn, dg = denoise(widen(x_hat), t_hat)
        if trace is not None:
            trace.append(dn.clone() if dg is None else torch.lerp(dg, dn, guidance))
        x_next = torch.empty_like(x_hat)
        _heun(lib, dn, dg, x_hat, d_cur, x_next, 0, guidance, t_hat, t_nxt)          # Euler step
        if i < num_steps - 1:
            dn, dg = denoise(widen(x_next), t_nxt)
            if trace is not None:
                trace.append(dn.clone() if dg is None else torch.lerp(dg, dn, guidance))
            _heun(lib, dn, dg, x_hat, d_cur, x_next, 1, guidance, t_hat, t_nxt)      # 2nd-order correction
    return x_next


class StackedRandomGenerator:
    """One torch.Generator per sample, seeded with seed % 2**32 (generate_images.py:120-134).
    Noise comes from torch so it is bit-identical to the reference's for the same seeds and device."""

    def __init__(self, device, seeds):
        self.generators = [torch.Generator(device).manual_seed(int(seed) % (1 << 32)) for seed in seeds]

    def randn(self, size, **kwargs):
        if size[0] != len(self.generators):
            raise AssertionError("size[0] must equal the number of seeds")
        return torch.stack([torch.randn(size[1:], generator=gen, **kwargs) for gen in self.generators])

    def randn_like(self, input):
        return self.randn(input.shape, dtype=input.dtype, layout=input.layout, device=input.device)

    def randint(self, *args, size, **kwargs):
        if size[0] != len(self.generators):
            raise AssertionError("size[0] must equal the number of seeds")
        return torch.stack([torch.randint(*args, size=size[1:], generator=gen, **kwargs) for gen in self.generators])
