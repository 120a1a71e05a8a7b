"""NVPrecond — drop-in for the reference's preconditioned denoiser object.

Same constructor arguments, attributes and call signature as the reference
(training/models.py:589-749; vanilla semantics: experiments/code/training/models.py:547-638),
so it can be passed as net / gnet / sr_model to generate_images_nvs and called as
`net(src, x, sigma, labels, conditioning_image)` by edm_sampler.  The forward runs entirely in
libvividb200.so: fp16 operands and residual stream on the tcgen05 path (fp32 accumulation and statistics), or — with
use_fp16=False / force_fp32=True, as in the reference (models.py:632,697) — the all-fp32 validation path (engine_f32.py).

Two semantics behind one class, chosen by the constructor arguments instead of
the reference's module-level VANILLA_MODE global:
  label_dim=...                               -> vanilla: B inputs, B outputs, geometry may be None
  source_label_dim=..., target_label_dim=...  -> dual-source: 2B interleaved inputs, B outputs
"""
import os

import torch

from . import _lib as L
from . import engine
from .networks import MPConv, MPFourier, SRXAttnUNet, UNetEncoder, XAttnUNet

_UNSUPPORTED = "is outside the B200 hot path (SURVEY.md §8(b)); there is no CPU fallback"


class NVPrecond(torch.nn.Module):
    def __init__(self, img_resolution, img_channels, label_dim=None, use_fp16=True, sigma_data=0.5, logvar_channels=128,
                 super_res=False, no_time_enc=None, depth_input=False, warp_depth_coor=False, uncond=None,
                 noisy_sr=0.25, source_label_dim=None, target_label_dim=None, **unet_kwargs):
        super().__init__()
        if depth_input or warp_depth_coor:
            raise NotImplementedError(f"depth_input / warp_depth_coor {_UNSUPPORTED}")
        if img_channels != 3:
            raise NotImplementedError("img_channels must be 3 (RGB pixel-space model)")
        if img_resolution < 4 or img_resolution & (img_resolution - 1):
            raise ValueError("img_resolution must be a power of two >= 4")
        self.dual = label_dim is None
        if self.dual:
            if source_label_dim is None or target_label_dim is None:
                raise TypeError("NVPrecond needs label_dim (vanilla) or source_label_dim+target_label_dim (dual-source)")
            if target_label_dim != 2 * source_label_dim:
                raise ValueError("dual-source mode expects target_label_dim == 2*source_label_dim")
            enc_label, unet_label = source_label_dim, target_label_dim
        else:
            enc_label = unet_label = label_dim
        self.init_kwargs = dict(img_resolution=img_resolution, img_channels=img_channels, use_fp16=use_fp16,
                                sigma_data=sigma_data, logvar_channels=logvar_channels, super_res=super_res,
                                no_time_enc=no_time_enc, depth_input=depth_input, warp_depth_coor=warp_depth_coor,
                                uncond=uncond, noisy_sr=noisy_sr, **unet_kwargs)
        self.init_kwargs.update(dict(source_label_dim=source_label_dim, target_label_dim=target_label_dim)
                                if self.dual else dict(label_dim=label_dim))
        self.init_args = ()
        self.img_resolution = img_resolution
        self.img_channels = img_channels
        self.label_dim = unet_label
        self.use_fp16 = use_fp16
        self.sigma_data = sigma_data
        self.super_res = super_res
        self.no_time_enc = no_time_enc
        self.depth_input = depth_input
        self.warp_depth_coor = warp_depth_coor
        self.uncond = uncond
        self.noisy_sr = noisy_sr
        self.encoder = UNetEncoder(img_resolution, img_channels, enc_label, **unet_kwargs) if not uncond else None
        make = SRXAttnUNet if super_res else XAttnUNet
        self.unet = make(img_resolution, img_channels, unet_label, **unet_kwargs)
        self.logvar_fourier = MPFourier(logvar_channels)
        self.logvar_linear = MPConv(logvar_channels, 1, kernel=[])
        self._plans = {}
        self.use_graph = True

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_reference(cls, ref):
        """Build from a reference NVPrecond instance (duck-typed: persistence re-creates classes from
        embedded source, torch_utils/persistence.py:226-237, so isinstance is useless)."""
        kw = dict(getattr(ref, "init_kwargs", {}))
        args = list(getattr(ref, "init_args", ()))
        names = ["img_resolution", "img_channels", "label_dim"]
        if "source_label_dim" in kw or (len(args) > 3):
            names = ["img_resolution", "img_channels", "source_label_dim", "target_label_dim"]
        for n, a in zip(names, args):
            kw[n] = a
        kw.pop("class_name", None)
        net = cls(**kw)
        sd = {k: v for k, v in ref.state_dict().items()}
        # persisted EMA snapshots are stored in fp16 (params and buffers): keep every tensor's dtype and bits
        own = dict(net.named_parameters())
        own.update(dict(net.named_buffers()))
        with torch.no_grad():
            for k, v in sd.items():
                if k in own and own[k].dtype != v.dtype:
                    own[k].data = own[k].data.to(v.dtype)
        net.load_state_dict(sd, strict=True)
        dev = next(ref.parameters()).device
        return net.to(dev).eval()

    def invalidate_plans(self):
        """Call after mutating weights (plans bake the normalised 16-bit weights)."""
        self._plans.clear()
        self._sig_tensors = None

    def _weight_signature(self):
        ps = getattr(self, "_sig_tensors", None)
        if ps is None:          # the parameter/buffer set of an inference net does not change: walk the module tree once
            ps = self._sig_tensors = list(self.parameters()) + list(self.buffers())
        return sum(t._version for t in ps)

    def plan(self, batch, device, fp32=False):
        """The recorded denoiser call for this batch size: the tcgen05 production plan, or — `fp32=True`, i.e.
        use_fp16=False / force_fp32=True in the reference's terms (models.py:632,697) — the fp32 validation plan."""
        key = (int(batch), str(device), bool(fp32))
        sig = self._weight_signature()
        hit = self._plans.get(key)
        if hit is not None and hit.weight_versions == sig:
            return hit
        if fp32:
            from .engine_f32 import PlanF32
            p = PlanF32(self, int(batch), device)
        elif os.environ.get("VB_LIB_PLAN") == "1":
            # the plan is recorded by the library itself (vb_net_plan_create) instead of engine.py: same ops, op for op
            # (tests/test_netplan.py) — the switch exists so that the whole parity suite can be run on that recorder
            from .netplan import LibPlan
            p = LibPlan(self, int(batch), device)
        else:
            p = engine.Plan(self, int(batch), device)
        p.weight_versions = sig
        self._plans[key] = p
        return p

    def _apply(self, fn, *a, **k):
        # .to()/.cuda()/.half() re-home the parameters the plans point at; a no-op call (generate_images_nvs does
        # net.to(device) on every invocation, generate_images.py:164-174) must not throw the plans, their tuning and
        # their captured graphs away
        before = [(q.data_ptr(), q.dtype, q.device) for q in self.parameters()]
        out = super()._apply(fn, *a, **k)
        if before != [(q.data_ptr(), q.dtype, q.device) for q in self.parameters()]:
            self._plans = {}
            self._sig_tensors = None
        return out

    def _logvar(self, sigma, B):
        """u(sigma) = logvar_linear(logvar_fourier(ln(sigma)/4)) as [B,1,1,1] fp32 (reference models.py:746-747; dual
        mode reads c_noise[::2], :686-688).  One small CUDA pass (vb_logvar), outside the replayed plan."""
        lv = torch.empty(B, dtype=torch.float32, device=sigma.device)
        w = self.logvar_linear.weight.detach().to(torch.float32).reshape(-1).contiguous()
        f = self.logvar_fourier.freqs.to(torch.float32).contiguous()
        ph = self.logvar_fourier.phases.to(torch.float32).contiguous()
        L.check(L.lib().vb_logvar(sigma.data_ptr(), B, 2 if self.dual else 1, w.data_ptr(), f.data_ptr(), ph.data_ptr(),
                                  w.numel(), lv.data_ptr(), torch.cuda.current_stream(sigma.device).cuda_stream), "vb_logvar")
        return lv.reshape(-1, 1, 1, 1)

    # ------------------------------------------------------------------ forward
    def forward(self, src, dst, sigma, geometry=None, conditioning_image=None, force_fp32=False, return_logvar=False,
                return_features=False, inject_features=None, **unet_kwargs):
        if (return_features or inject_features is not None) and (self.encoder is None or self.super_res):
            raise NotImplementedError(f"return_features / inject_features on a net without source-view encoder {_UNSUPPORTED}")
        fp32 = bool(force_fp32) or not self.use_fp16
        if fp32 and (return_features or inject_features is not None):
            raise NotImplementedError("return_features / inject_features are not offered by the fp32 validation path")
        if unet_kwargs:
            raise TypeError(f"unexpected arguments {sorted(unet_kwargs)}")
        if dst.device.type != "cuda":
            raise RuntimeError("vivid_b200.NVPrecond runs on CUDA (sm_100a) only; there is no CPU fallback")
        R = self.img_resolution
        if dst.ndim != 4 or dst.shape[1:] != (3, R, R):
            raise ValueError(f"dst must be [N,3,{R},{R}], got {tuple(dst.shape)}")
        n_in = dst.shape[0]
        if self.dual:
            if n_in % 2:
                raise ValueError("dual-source mode expects 2B interleaved inputs")
            if geometry is None:
                raise TypeError("dual-source mode requires geometry")   # reference: NoneType * int (models.py:631)
        B = n_in // 2 if self.dual else n_in
        with torch.no_grad(), torch.cuda.device(dst.device):
            p = self.plan(B, dst.device, fp32=fp32)
            p.in_x.copy_(dst)
            if inject_features is not None:
                # no_time_enc caching (edm_sampler, generate_images.py:52-57): the encoder is skipped and the cached maps
                # feed the cross-attention blocks; src is not read (reference: training/models.py:664-665)
                views = p.feature_views()
                if len(inject_features) != len(views):
                    raise ValueError(f"inject_features must hold {len(views)} feature maps, got {len(inject_features)}")
                for v, f in zip(views, inject_features):
                    if f.shape != v.shape:
                        raise ValueError(f"feature map must be {tuple(v.shape)}, got {tuple(f.shape)}")
                    v.copy_(f)
            elif p.in_src is not None:
                if src.shape != p.in_src.shape:
                    raise ValueError(f"src must be {tuple(p.in_src.shape)}, got {tuple(src.shape)}")
                p.in_src.copy_(src)
            sig = torch.as_tensor(sigma, dtype=torch.float32, device=dst.device).reshape(-1)
            p.in_sigma.copy_(sig.expand(n_in) if sig.numel() == 1 else sig)
            if geometry is None:
                p.in_geom.zero_()
            else:
                g = geometry.to(torch.float32).reshape(-1, p.in_geom.shape[1])
                p.in_geom.copy_(g.expand(n_in, -1) if g.shape[0] == 1 else g)
            if self.super_res:
                if conditioning_image is None:
                    raise AssertionError("super_res model requires conditioning_image")
                p.in_cond.copy_(conditioning_image)
                # the reference draws this from the GLOBAL torch RNG on every call;
                # same call, same device/dtype/shape, so identical noise for identical generator state
                p.in_noise.copy_(torch.randn_like(conditioning_image))
            if return_features:
                # the reference returns before the denoising UNet runs (training/models.py:671-672)
                p.run(graph=self.use_graph, section="enc")
                return [v.clone(memory_format=torch.channels_last) for v in p.feature_views()]
            p.run(graph=self.use_graph, section="all" if inject_features is None else "unet")
            if return_logvar:
                return p.out_d.clone(), self._logvar(p.in_sigma, B)
            return p.out_d.clone()
