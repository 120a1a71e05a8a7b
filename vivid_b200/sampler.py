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


def _heun(lib, d_net, d_gnet, x_hat, d_cur, x_next, phase, guidance, t_hat, t_next):
    desc = L.HeunDesc(d_net=d_net.data_ptr(), d_gnet=L.ptr(d_gnet), x_hat=x_hat.data_ptr(), d_cur=d_cur.data_ptr(),
                      x_next=x_next.data_ptr(), n=x_hat.numel(), phase=phase, guidance=float(guidance),
                      t_hat=float(t_hat), t_next=float(t_next))
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
                dtype=torch.float32, randn_like=torch.randn_like):
    if dtype != torch.float32:
        raise NotImplementedError("the sampler state is fp32 (the reference default); other dtypes are not implemented")
    if noise.device.type != "cuda":
        raise RuntimeError("vivid_b200.edm_sampler runs on CUDA only; there is no CPU fallback")
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
        x_next = torch.empty_like(x_hat)
        _heun(lib, dn, dg, x_hat, d_cur, x_next, 0, guidance, t_hat, t_nxt)          # Euler step
        if i < num_steps - 1:
            dn, dg = denoise(widen(x_next), t_nxt)
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
