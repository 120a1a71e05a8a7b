"""fp32 validation engine — NVPrecond(use_fp16=False) / forward(force_fp32=True).

The reference switches its whole forward to fp32 with `use_fp16=False` or `force_fp32=True`
(training/models.py:632,697).  The production plan (engine.Plan) stores activations in fp16 like the reference's
reduced-precision mode and runs the GEMMs on tcgen05; this plan runs the same network op by op, unfused, in fp32 on the
CUDA cores (csrc/f32path.cu) so that the north star's fp32-mode bound — denoiser rel-L2 <= 1e-4 against the
reference's fp32 path — can be checked on real weights.  It follows UNet.forward / Block.forward literally
(training/models.py:165-206, 251-315, 385-414, 483-518, 536-570); it is for validation, not speed, replays eagerly
(no CUDA graph) and does not pool buffers.
"""
import ctypes as C
import math

import torch

from . import _lib as L
from . import engine


class PlanF32(engine.Plan):
    # ------------------------------------------------------------------ op emitters (each records one eager launch)
    def f32(self, rows, ch):
        return self.buf((rows, ch), torch.float32)

    def _emit(self, fn, desc, what):
        self.descs.append(desc)
        self.ops.append(lambda: L.check(fn(C.byref(desc), self.stream), what))

    def conv32(self, x, w, B, R, cin, cout, taps, ldo=None):
        out = self.f32(B * R * R, ldo or cout)
        d = L.F32ConvDesc(x=x.data_ptr(), w=w.data_ptr(), out=out.data_ptr(), B=B, H=R, W=R, cin=cin, cout=cout, taps=taps,
                          ldo=ldo or cout)
        self._emit(self.lib.vb_f32_conv, d, "vb_f32_conv")
        self.alg_flops += 2.0 * B * R * R * cin * cout * taps
        return out

    def op32(self, kind, a, B, R, ca, out, **kw):
        d = L.F32OpDesc(a=a.data_ptr(), out=out.data_ptr(), kind=kind, B=B, H=R, W=R, ca=ca)
        for k, v in kw.items():
            if isinstance(v, torch.Tensor):
                v = v.data_ptr()
            if k in ("part_seq", "part_off"):
                v = (L.i32 * 3)(*v)
            setattr(d, k, v)
        self._emit(self.lib.vb_f32_op, d, "vb_f32_op")
        return out

    def act32(self, a, B, R, ch, flags, mod=None, mod_stride=0):
        return self.op32(L.VB_F32_ACT, a, B, R, ch, self.f32(B * R * R, ch), flags=flags, mod=mod, mod_stride=mod_stride)

    def sum32(self, a, b, t, B, R, ch, clip=None):
        n = math.sqrt((1 - t) ** 2 + t ** 2)
        return self.op32(L.VB_F32_SUM, a, B, R, ch, self.f32(B * R * R, ch), b=b, wa=(1 - t) / n, wb=t / n,
                         clip=float(clip) if clip is not None else 0.0)

    def w32(self, w, gain=1.0, perm=(0, 0)):
        """vb_weight_prep to fp32 [cout][cin*taps] (forced weight normalisation in fp32, models.py:115-121)."""
        w = w.detach().contiguous()
        cout, cin = w.shape[0], w.shape[1]
        taps = w.shape[2] * w.shape[3] if w.ndim == 4 else 1
        dst = self.buf((cout, cin * taps), torch.float32)
        d = L.WeightPrepDesc(src=w.data_ptr(), dst=dst.data_ptr(), src_dtype=engine._DT[w.dtype], dst_dtype=L.VB_F32, cout=cout,
                             cin=cin, taps=taps, cout_pad=cout, split=cin, seg_a_pad=cin, seg_b_pad=0, perm_parts=perm[0],
                             perm_dim=perm[1], gain=float(gain), scale_a=1.0, scale_b=1.0)
        self.keep.append(w)
        L.check(self.lib.vb_weight_prep(C.byref(d), self.stream), "vb_weight_prep")
        return dst

    def embed32(self, *a, **k):
        mod, offs, total = self.embed(*a, **k)                 # fp32 already; recorded in the C plan
        idx = self.lib.vb_plan_num_ops(self.handle) - 1
        self.ops.append(lambda: L.check(self.lib.vb_plan_run(self.handle, idx, idx + 1, self.stream), "vb_plan_run"))
        return mod, offs, total

    # ------------------------------------------------------------------ one UNet / encoder (models.py:385-414,536-570)
    def run_unet32(self, unet, x, B, mod, offs, mod_total, features=None, feat_seg=1, zero_feature_keys=False,
                   collect_features=False):
        R = unet.img_resolution
        C_cur = None
        skips, feats_out = [], []
        features = list(features or [])
        cur = x
        for s in unet.enc_specs + unet.dec_specs:
            m = unet.enc[s.name] if s.group == "enc" else unet.dec[s.name]
            if s.kind == "conv":
                cur = self.conv32(cur, self.w32(m.weight), B, R, s.cin, s.cout, 9)
                C_cur = s.cout
                skips.append((cur, s.cout))
                continue
            xx = cur
            if s.group == "dec" and s.skip_ch:
                skip, sc = skips.pop()
                wa, wb = self.cat_weights(C_cur, sc, unet.concat_balance)
                xx = self.op32(L.VB_F32_CAT, xx, B, R, C_cur, self.f32(B * R * R, C_cur + sc), b=skip, cb=sc, wa=wa, wb=wb)
                C_cur += sc
            assert C_cur == s.cin, (s.name, C_cur, s.cin)
            if s.resample != "keep":
                R = s.res
                xx = self.op32(L.VB_F32_DOWN if s.resample == "down" else L.VB_F32_UP, xx, B, R, C_cur,
                               self.f32(B * R * R, C_cur))
            Cc = s.cout
            if s.flavor == "enc":
                if m.conv_skip is not None:
                    xx = self.conv32(xx, self.w32(m.conv_skip.weight), B, R, s.cin, Cc, 1)
                xx = self.act32(xx, B, R, Cc, L.VB_F32_NORM)
            c0 = Cc if s.flavor == "enc" else s.cin
            y = self.act32(xx, B, R, c0, L.VB_F32_SILU)
            y = self.conv32(y, self.w32(m.conv_res0.weight), B, R, c0, Cc, 9)
            y = self.act32(y, B, R, Cc, L.VB_F32_MOD | L.VB_F32_SILU, mod=mod[:, offs[(s.group, s.name)]:], mod_stride=mod_total)
            y = self.conv32(y, self.w32(m.conv_res1.weight), B, R, Cc, Cc, 9)
            if s.flavor == "dec" and m.conv_skip is not None:
                xx = self.conv32(xx, self.w32(m.conv_skip.weight), B, R, s.cin, Cc, 1)
            if s.heads == 0:
                xx = self.sum32(xx, y, m.res_balance, B, R, Cc, clip=m.clip_act)
            else:
                xx = self.sum32(xx, y, m.res_balance, B, R, Cc)
                S, D, h = R * R, s.head_dim, s.heads
                nseg = feat_seg if s.xattn else 0
                real_seg = 0 if zero_feature_keys else nseg
                sk = S * (1 + real_seg)
                q = self.f32(B * h * S, D)
                k = self.f32(B * h * sk, D)
                v = self.f32(B * h * sk, D)
                qkv = self.conv32(xx, self.w32(m.attn_qkv.weight, perm=(3, D)), B, R, Cc, 3 * Cc, 1)
                self.op32(L.VB_F32_QKV, qkv, B, R, 3 * Cc, q, out2=k, out3=v, heads=h, parts=3, head_dim=D, seg_div=1,
                          part_seq=(S, sk, sk), part_off=(0, 0, 0))
                if s.xattn and not zero_feature_keys:
                    f, fB = features.pop(0)
                    kv = self.conv32(f, self.w32(m.x_attn_kv.weight, perm=(2, D)), fB, R, Cc, 2 * Cc, 1)
                    self.op32(L.VB_F32_QKV, kv, fB, R, 2 * Cc, k, out2=v, heads=h, parts=2, head_dim=D, seg_div=feat_seg,
                              part_seq=(sk, sk, 0), part_off=(S, S, 0))
                y = self.f32(B * S, Cc)
                args = (q.data_ptr(), k.data_ptr(), v.data_ptr(), y.data_ptr(), B, h, S, sk, D,
                        S * nseg if zero_feature_keys else 0)
                self.ops.append(lambda a=args: L.check(self.lib.vb_f32_attn(*a, self.stream), "vb_f32_attn"))
                self.alg_flops += 4.0 * B * h * S * sk * D
                y = self.conv32(y, self.w32(m.attn_proj.weight), B, R, Cc, Cc, 1)
                xx = self.sum32(xx, y, m.attn_balance, B, R, Cc, clip=m.clip_act)
            if collect_features and s.heads > 0:
                feats_out.append((xx, B))
            if s.group == "enc":
                skips.append((xx, Cc))
            cur, C_cur = xx, Cc
        raw = None
        if unet.out_conv is not None:
            wo = self.w32(unet.out_conv.weight, gain=float(unet.out_gain.detach().float().item()))
            raw = self.conv32(cur, wo, B, R, C_cur, unet.out_conv.out_channels, 9, ldo=4)
        return raw, feats_out

    # ------------------------------------------------------------------ whole NVPrecond call
    def _build(self):
        net, B, Bx = self.net, self.B, self.Bx
        R = net.img_resolution
        sd = float(net.sigma_data)
        self.ops, self.descs = [], []
        self.in_x = self.buf((Bx, 3, R, R), torch.float32, zero=True)
        self.in_src = self.buf((Bx, 3, R, R), torch.float32, zero=True) if net.encoder is not None else None
        self.in_sigma = self.buf((Bx,), torch.float32)
        self.in_sigma.fill_(1.0)
        ldim_enc = net.encoder.label_dim if net.encoder is not None else 0
        ldim_unet = net.unet.label_dim
        self.in_geom = self.buf((Bx, max(ldim_enc, ldim_unet // (2 if self.dual else 1), 1)), torch.float32, zero=True)
        self.in_cond = self.buf((B, 3, R, R), torch.float32, zero=True) if net.super_res else None
        self.in_noise = self.buf((B, 3, R, R), torch.float32, zero=True) if net.super_res else None
        self.out_d = self.buf((B, 3, R, R), torch.float32, zero=True)
        geom_scale = 0.0 if net.uncond else 1.0

        features, feat_seg = None, 1
        if net.encoder is not None:
            enc = net.encoder
            src = self.op32(L.VB_F32_PRECOND_IN, self.in_src, Bx, R, 4, self.f32(Bx * R * R, 4), img_stride=3 * R * R)
            mod, offs, total = self.embed32(enc, Bx, self.in_sigma, 1, self.in_geom if ldim_enc else None, Bx, ldim_enc,
                                            0.0 if net.no_time_enc else 1.0, geom_scale)
            _, features = self.run_unet32(enc, src, Bx, mod, offs, total, collect_features=True)
            feat_seg = 2 if self.dual else 1
        self.enc_ops, self.features = 0, []

        unet = net.unet
        step = 2 if self.dual else 1
        cin = 7 if net.super_res else 4
        x = self.op32(L.VB_F32_PRECOND_IN, self.in_x, B, R, cin, self.f32(B * R * R, cin), img_stride=3 * R * R * step,
                      mod=self.in_sigma, mod_stride=step, b=self.in_cond, b2=self.in_noise, wa=sd,
                      wb=float(net.noisy_sr if net.noisy_sr is not None else 0.0))
        mod, offs, total = self.embed32(unet, B, self.in_sigma, step, self.in_geom if ldim_unet else None, B, ldim_unet, 1.0,
                                        geom_scale)
        raw, _ = self.run_unet32(unet, x, B, mod, offs, total, features=features, feat_seg=feat_seg,
                                 zero_feature_keys=net.encoder is None)
        d = L.PrecondOutDesc(x=self.in_x.data_ptr(), f=raw.data_ptr(), sigma=self.in_sigma.data_ptr(), d_out=self.out_d.data_ptr(),
                             B=B, R=R, ldf=4, sigma_n=B, sigma_stride=step, img_stride=3 * R * R * step, sigma_data=sd)
        self._emit(self.lib.vb_precond_out, d, "vb_precond_out")
        self.num_ops = len(self.ops)
        self.launches = len(self.ops)
        self.padded_flops = self.alg_flops

    def run(self, graph=False, section="all"):
        if section != "all":
            raise NotImplementedError("return_features / inject_features are not offered by the fp32 validation path")
        self.stream = torch.cuda.current_stream(self.device).cuda_stream
        for op in self.ops:
            op()

    def profile(self, repeats=3):
        raise NotImplementedError("the fp32 validation path is not a measured path")
