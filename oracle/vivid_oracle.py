"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not shipped, not measured, never on the product path.

A plain-PyTorch (CPU or GPU, fp32) restatement of the reference's guided EDM2 denoising
path, written functionally over a state_dict so that it shares no code with the product
package `vivid_b200`.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it.

Each function cites the reference lines it restates.  Two semantics are covered:
  * "vanilla" — snapshot tree  /root/reference/experiments/code/training/models.py
                (the only tree in which guidance / uncond gnet / SR run)
  * "dual"    — current tree   /root/reference/training/models.py (two source views per target)

Parity pinning: the reference has no tests or golden vectors, so this oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by
tests/golden/make_golden.py (denoiser / sampler vectors of both trees), make_golden_extra.py (return_logvar,
return_features / inject_features with no_time_enc, S_churn > 0) and make_golden_metrics.py (the statistics leg of
calculate_metrics.py gen, run through the reference's own code with a fake detector) — all import /root/reference —
and committed under tests/golden/.  tests/test_oracle_golden.py replays them.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------- primitive ops


def normalize(x, dim=None, eps=1e-4):
    """training/models.py:37-42 — x / (eps + ||x||_2 / sqrt(n_reduced)), statistic in fp32."""
    if dim is None:
        dim = list(range(1, x.ndim))
    n = torch.linalg.vector_norm(x, dim=dim, keepdim=True, dtype=torch.float32)
    n = eps + n * math.sqrt(n.numel() / x.numel())
    return x / n.to(x.dtype)


def resample(x, mode):
    """training/models.py:48-61 with f=[1,1]: 'down' is a 2x2 mean pool, 'up' is nearest x2."""
    if mode == "keep":
        return x
    if mode == "down":
        return F.avg_pool2d(x, 2)
    assert mode == "up"
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def mp_silu(x):
    """training/models.py:66-67."""
    return F.silu(x) / 0.596


def mp_sum(a, b, t=0.5):
    """training/models.py:72-73."""
    return a.lerp(b, t) / math.sqrt((1 - t) ** 2 + t ** 2)


def mp_cat(a, b, t=0.5):
    """training/models.py:78-84 (dim=1)."""
    na, nb = a.shape[1], b.shape[1]
    c = math.sqrt((na + nb) / ((1 - t) ** 2 + t ** 2))
    return torch.cat([a * (c / math.sqrt(na) * (1 - t)), b * (c / math.sqrt(nb) * t)], dim=1)


def mp_fourier(x, freqs, phases):
    """MPFourier.forward, training/models.py:96-101."""
    y = x.to(torch.float32).ger(freqs.to(torch.float32)) + phases.to(torch.float32)
    return (y.cos() * math.sqrt(2)).to(x.dtype)


def mp_conv(x, w, gain=1.0):
    """MPConv.forward (eval mode), training/models.py:114-126."""
    w = normalize(w.to(torch.float32))
    w = w * (float(gain) / math.sqrt(w[0].numel()))
    w = w.to(x.dtype)
    if w.ndim == 2:
        return x @ w.t()
    return F.conv2d(x, w, padding=w.shape[-1] // 2)


# ----------------------------------------------------------------------------- state-dict views


class Params:
    """Ordered view of a (sub-)state_dict; `p.sub('unet.')` strips a prefix."""

    def __init__(self, sd):
        self.sd = dict(sd)

    def sub(self, prefix):
        return Params({k[len(prefix):]: v for k, v in self.sd.items() if k.startswith(prefix)})

    def has(self, key):
        return key in self.sd

    def __getitem__(self, key):
        return self.sd[key]

    def blocks(self, group):
        """Ordered block names of ModuleDict `group` ('enc' or 'dec') — registration order."""
        names = []
        for k in self.sd:
            if k.startswith(group + "."):
                n = k[len(group) + 1:].split(".")[0]
                if n not in names:
                    names.append(n)
        return names


def _attention(q, k, v):
    """softmax(q.k/sqrt(D)) v over keys; q [B,h,D,Sq], k/v [B,h,D,Sk]
    (snapshot models.py:190-191 einsum form == current tree :198/:305 SDPA form)."""
    w = torch.einsum("nhcq,nhck->nhqk", q, k / math.sqrt(q.shape[2])).softmax(dim=3)
    return torch.einsum("nhqk,nhck->nhcq", w, v)


def block_forward(p, x, emb, flavor, mode, heads, features=(), res_balance=0.3, attn_balance=0.3, clip_act=256):
    """Block.forward / XAttnBlock.forward — training/models.py:165-206, 251-315
    (snapshot: experiments/code/training/models.py:162-198, 242-287).
    `features` holds 0 (self-attention only), 1 (vanilla) or 2 (dual-source) cross feature maps."""
    x = resample(x, mode)
    if flavor == "enc":
        if p.has("conv_skip.weight"):
            x = mp_conv(x, p["conv_skip.weight"])
        x = normalize(x, dim=1)
    y = mp_conv(mp_silu(x), p["conv_res0.weight"])
    c = mp_conv(emb, p["emb_linear.weight"], gain=p["emb_gain"]) + 1
    y = mp_silu(y * c.unsqueeze(2).unsqueeze(3).to(y.dtype))
    y = mp_conv(y, p["conv_res1.weight"])
    if flavor == "dec" and p.has("conv_skip.weight"):
        x = mp_conv(x, p["conv_skip.weight"])
    x = mp_sum(x, y, t=res_balance)
    if heads:
        B, Cc, H, W = x.shape
        qkv = mp_conv(x, p["attn_qkv.weight"]).reshape(B, heads, -1, 3, H * W)
        q, k, v = normalize(qkv, dim=2).unbind(3)
        ks, vs = [k], [v]
        for f in features:
            kv = mp_conv(f, p["x_attn_kv.weight"]).reshape(B, heads, -1, 2, f.shape[2] * f.shape[3])
            xk, xv = normalize(kv, dim=2).unbind(3)
            ks.append(xk)
            vs.append(xv)
        y = _attention(q, torch.cat(ks, -1), torch.cat(vs, -1))
        y = mp_conv(y.reshape(B, Cc, H, W), p["attn_proj.weight"])
        x = mp_sum(x, y, t=attn_balance)
    if clip_act is not None:
        x = x.clip(-clip_act, clip_act)
    return x


def _embedding(p, noise_labels, geometry, label_balance):
    """UNet.forward embedding, training/models.py:388-391 (snapshot :361-364)."""
    emb = mp_conv(mp_fourier(noise_labels, p["emb_fourier.freqs"], p["emb_fourier.phases"]), p["emb_noise.weight"])
    if p.has("emb_label.weight") and geometry is not None:
        emb = mp_sum(emb, mp_conv(geometry, p["emb_label.weight"]), t=label_balance)
    return mp_silu(emb)


def _heads(p, name, group, channels_per_head):
    key = f"{group}.{name}.attn_qkv.weight"
    if not p.has(key):
        return 0
    return (p[key].shape[0] // 3) // channels_per_head


def _mode(name):
    return "down" if name.endswith("_down") else "up" if name.endswith("_up") else "keep"


def encoder_forward(p, x, noise_labels, geometry, label_balance=0.5, concat_balance=0.5):
    """UNetEncoder.forward, training/models.py:536-570 (snapshot :501-527): returns the output of
    every attention block; trailing non-attention decoder blocks do not exist in the state."""
    emb = _embedding(p, noise_labels, geometry, label_balance)
    x = torch.cat([x, torch.ones_like(x[:, :1])], dim=1)
    skips, feats = [], []
    for name in p.blocks("enc"):
        if "conv" in name:
            x = mp_conv(x, p[f"enc.{name}.weight"])
        else:
            h = _heads(p, name, "enc", 64)
            x = block_forward(p.sub(f"enc.{name}."), x, emb, "enc", _mode(name), h)
            if h > 0:
                feats.append(x)
        skips.append(x)
    for name in p.blocks("dec"):
        if "block" in name:
            x = mp_cat(x, skips.pop(), t=concat_balance)
        h = _heads(p, name, "dec", 64)
        x = block_forward(p.sub(f"dec.{name}."), x, emb, "dec", _mode(name), h)
        if h > 0:
            feats.append(x)
    return feats


def unet_forward(p, x, features, noise_labels, geometry, channels_per_head=64, dual=False, label_balance=0.5,
                 concat_balance=0.5):
    """XAttnUNet.forward — vanilla: snapshot models.py:455-483; dual: training/models.py:483-518
    (features are 2B-interleaved: [0::2] first source, [1::2] second source)."""
    emb = _embedding(p, noise_labels, geometry, label_balance)
    features = list(features)

    def pop():
        f = features.pop(0)
        return (f[0::2], f[1::2]) if dual else (f,)

    x = torch.cat([x, torch.ones_like(x[:, :1])], dim=1)
    skips = []
    for name in p.blocks("enc"):
        if "conv" in name:
            x = mp_conv(x, p[f"enc.{name}.weight"])
        else:
            bp = p.sub(f"enc.{name}.")
            h = _heads(p, name, "enc", channels_per_head)
            fs = pop() if bp.has("x_attn_kv.weight") else ()
            x = block_forward(bp, x, emb, "enc", _mode(name), h, fs)
        skips.append(x)
    for name in p.blocks("dec"):
        if "block" in name:
            x = mp_cat(x, skips.pop(), t=concat_balance)
        bp = p.sub(f"dec.{name}.")
        h = _heads(p, name, "dec", channels_per_head)
        fs = pop() if bp.has("x_attn_kv.weight") else ()
        x = block_forward(bp, x, emb, "dec", _mode(name), h, fs)
    return mp_conv(x, p["out_conv.weight"], gain=p["out_gain"])


# ----------------------------------------------------------------------------- NVPrecond


class OracleNet:
    """Functional NVPrecond: `net(src, dst, sigma, geometry, conditioning_image)`.

    cfg keys: img_resolution, img_channels, label_dim (vanilla) | source_label_dim/target_label_dim (dual),
    super_res, uncond, noisy_sr, sigma_data, no_time_enc, dual.
    """

    def __init__(self, state_dict, cfg):
        self.p = Params({k: v.to(torch.float32) for k, v in state_dict.items()})
        self.cfg = dict(cfg)
        self.img_resolution = cfg["img_resolution"]
        self.img_channels = cfg.get("img_channels", 3)
        self.super_res = bool(cfg.get("super_res", False))
        self.uncond = bool(cfg.get("uncond", False))
        self.noisy_sr = cfg.get("noisy_sr", 0.25)
        self.sigma_data = cfg.get("sigma_data", 0.5)
        self.no_time_enc = cfg.get("no_time_enc", None)
        self.dual = bool(cfg.get("dual", False))
        self.label_dim = cfg.get("label_dim", cfg.get("target_label_dim", 0))
        self.depth_input = False

    def to(self, device):
        self.p = Params({k: v.to(device) for k, v in self.p.sd.items()})
        return self

    def xattn_feature_shapes(self):
        """(channels, resolution) of every cross-attention block in consumption order."""
        u = self.p.sub("unet.")
        out = []
        for grp in ("enc", "dec"):
            for name in u.blocks(grp):
                if u.has(f"{grp}.{name}.x_attn_kv.weight"):
                    out.append((u[f"{grp}.{name}.x_attn_kv.weight"].shape[1], int(name.split("x")[0])))
        return out

    def __call__(self, src, dst, sigma, geometry=None, conditioning_image=None, return_features=False,
                 inject_features=None, sr_noise=None, return_logvar=False):
        if self.dual:
            d = self._forward_dual(src, dst, sigma, geometry, conditioning_image, return_features, inject_features)
        else:
            d = self._forward_vanilla(src, dst, sigma, geometry, conditioning_image, return_features, inject_features,
                                      sr_noise)
        if return_logvar and not return_features:
            return d, self.logvar(sigma)
        return d

    def logvar(self, sigma):
        """Uncertainty head: snapshot training/models.py:746-747; current tree :686-688 reads c_noise[::2]."""
        sigma = torch.as_tensor(sigma, dtype=torch.float32, device=self.p["logvar_linear.weight"].device).reshape(-1)
        c_noise = self._coeffs(sigma)[3]
        if self.dual:
            c_noise = c_noise[::2]
        feat = mp_fourier(c_noise, self.p["logvar_fourier.freqs"], self.p["logvar_fourier.phases"])
        return mp_conv(feat, self.p["logvar_linear.weight"]).reshape(-1, 1, 1, 1)

    def _coeffs(self, sigma):
        sd = self.sigma_data
        c_skip = sd ** 2 / (sigma ** 2 + sd ** 2)
        c_out = sigma * sd / (sigma ** 2 + sd ** 2).sqrt()
        c_in = 1 / (sd ** 2 + sigma ** 2).sqrt()
        c_noise = sigma.flatten().log() / 4
        return c_skip, c_out, c_in, c_noise

    def _forward_vanilla(self, src, dst, sigma, geometry, cond, return_features, inject_features, sr_noise):
        """snapshot experiments/code/training/models.py:581-638."""
        x = dst.to(torch.float32)
        sigma = torch.as_tensor(sigma, dtype=torch.float32, device=x.device).reshape(-1, 1, 1, 1)
        if self.label_dim == 0:
            geometry = None
        elif geometry is None:
            geometry = torch.zeros([1, self.label_dim], device=x.device)
        else:
            geometry = geometry.to(torch.float32).reshape(-1, self.label_dim)
        if geometry is not None:
            geometry = geometry * int(not self.uncond)
        c_skip, c_out, c_in, c_noise = self._coeffs(sigma)
        x_in = c_in * x
        if self.super_res:
            assert cond is not None
            noise = torch.randn_like(cond) if sr_noise is None else sr_noise   # global RNG
            x_in = torch.cat([x_in, cond + self.noisy_sr * noise], dim=1)
        if inject_features is not None:
            features = [f.clone() for f in inject_features]
        elif self.uncond:
            features = [torch.zeros((x_in.shape[0], c, r, r), dtype=x_in.dtype, device=x_in.device)
                        for c, r in self.xattn_feature_shapes()]
        else:
            features = encoder_forward(self.p.sub("encoder."), src.to(torch.float32),
                                       c_noise * int(not self.no_time_enc), geometry)
        if return_features:
            return features
        f_x = unet_forward(self.p.sub("unet."), x_in, features, c_noise, geometry,
                           channels_per_head=32 if self.super_res else 64)
        return c_skip * dst + c_out * f_x.to(torch.float32)

    def _forward_dual(self, src, dst, sigma, geometry, cond, return_features, inject_features):
        """current tree training/models.py:628-689 (_forward_dualsource): 2B interleaved inputs, B outputs."""
        x = dst.to(torch.float32)
        sigma = torch.as_tensor(sigma, dtype=torch.float32, device=x.device).reshape(-1, 1, 1, 1)
        geometry = geometry.to(torch.float32) * int(not self.uncond)
        c_skip, c_out, c_in, c_noise = self._coeffs(sigma)
        x_in = c_in * x
        if self.super_res:
            noise = torch.randn_like(cond)
            x_in = torch.cat([x_in, (cond + self.noisy_sr * noise).repeat_interleave(2, dim=0)], dim=1)
        if inject_features is not None:
            features = [f.clone() for f in inject_features]
        else:
            features = encoder_forward(self.p.sub("encoder."), src.to(torch.float32),
                                       c_noise * int(not self.no_time_enc), geometry)
        if return_features:
            return features
        b = x_in.shape[0] // 2
        f_x = unet_forward(self.p.sub("unet."), x_in[::2], features, c_noise[::2], geometry.reshape(b, -1),
                           channels_per_head=32 if self.super_res else 64, dual=True)
        return c_skip[::2] * dst[::2] + c_out[::2] * f_x.to(torch.float32)


# ----------------------------------------------------------------------------- sampler


def sigma_schedule(num_steps, sigma_min, sigma_max, rho, device, dtype=torch.float32):
    """generate_images.py:68-70 — rho schedule plus t_N = 0."""
    i = torch.arange(num_steps, dtype=dtype, device=device)
    t = (sigma_max ** (1 / rho) + i / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    return torch.cat([t, torch.zeros_like(t[:1])])


def edm_sampler(net, src, noise, labels=None, gnet=None, conditioning_image=None, num_steps=32, sigma_min=0.002,
                sigma_max=80, rho=7, guidance=1, S_churn=0, S_min=0, S_max=float("inf"), S_noise=1,
                dtype=torch.float32, randn_like=torch.randn_like, trace=None):
    """EDM Heun sampler with autoguidance and the stochastic churn branch (generate_images.py:77-84).
    vanilla: snapshot generate_images.py:41-91; dual-source fold: current generate_images.py:43-118."""
    features = None
    if getattr(net, "no_time_enc", None):
        features = net(src, torch.zeros_like(src), torch.ones(src.shape[0], dtype=dtype, device=noise.device), labels,
                       conditioning_image, return_features=True)

    def denoise(x, t):
        if getattr(net, "dual", False):
            t = t.expand(x.shape[0])
        dx = net(src, x, t, labels, conditioning_image, inject_features=features).to(dtype)
        if guidance == 1:
            return dx
        ref = gnet(src, x, t).to(dtype)
        return ref.lerp(dx, guidance)

    t_steps = sigma_schedule(num_steps, sigma_min, sigma_max, rho, noise.device, dtype)
    x_next = noise.to(dtype) * t_steps[0]
    dual = False
    for i, (t_cur, t_next) in enumerate(zip(t_steps[:-1], t_steps[1:])):
        x_cur = x_next
        if S_churn > 0 and S_min <= t_cur <= S_max:
            gamma = min(S_churn / num_steps, math.sqrt(2) - 1)
            t_hat = t_cur + gamma * t_cur
            x_hat = x_cur + (t_hat ** 2 - t_cur ** 2).sqrt() * S_noise * randn_like(x_cur)
        else:
            x_hat, t_hat = x_cur, t_cur
        d0 = denoise(x_hat, t_hat)
        if trace is not None:
            trace.append(d0.clone())
        dual = d0.shape[0] != x_hat.shape[0]
        if dual:
            d_cur = (x_hat[::2] - d0) / t_hat
            half = x_hat[::2] + (t_next - t_hat) * d_cur
            x_next = half.repeat_interleave(2, dim=0)
        else:
            d_cur = (x_hat - d0) / t_hat
            x_next = x_hat + (t_next - t_hat) * d_cur
        if i < num_steps - 1:
            d1 = denoise(x_next, t_next)
            if trace is not None:
                trace.append(d1.clone())
            if dual:
                d_prime = (x_next[::2] - d1) / t_next
                half = x_hat[::2] + (t_next - t_hat) * (0.5 * d_cur + 0.5 * d_prime)
                x_next = half.repeat_interleave(2, dim=0)
            else:
                d_prime = (x_next - d1) / t_next
                x_next = x_hat + (t_next - t_hat) * (0.5 * d_cur + 0.5 * d_prime)
    return x_next[::2] if dual else x_next


class StackedRandomGenerator:
    """generate_images.py:120-134 — one torch.Generator per sample, seeded seed % 2**32."""

    def __init__(self, device, seeds):
        self.generators = [torch.Generator(device).manual_seed(int(s) % (1 << 32)) for s in seeds]

    def randn(self, size, **kw):
        assert size[0] == len(self.generators)
        return torch.stack([torch.randn(size[1:], generator=g, **kw) for g in self.generators])


def encode_latents(x):
    """training/encoders.py:58-59."""
    return x.to(torch.float32) / 127.5 - 1


def decode(x):
    """training/encoders.py:61-62."""
    return (x.to(torch.float32) * 127.5 + 128).clip(0, 255).to(torch.uint8)


# Geometry packing statistics, training/utils.py:38-44 (values are data, restated verbatim).
GEOM_MEAN = [9.6681e-01, -1.6038e-04, -3.7034e-05, -1.6904e-03, -8.7718e-05, 9.9869e-01, 3.1288e-03, -1.0794e-03,
             1.0653e-05, 3.0997e-03, 9.6691e-01, 1.2561e-02, 5.7708e+01, 5.7704e+01, 3.2000e+01, 3.2000e+01,
             5.7708e+01, 5.7704e+01, 3.2000e+01, 3.2000e+01]
GEOM_STD = [0.1104, 0.0346, 0.2279, 0.4930, 0.0347, 0.0091, 0.0367, 0.2208, 0.2279, 0.0368, 0.1088, 1.0751, 6.6464,
            6.6511, 0.0, 0.0, 6.6464, 6.6511, 0.0, 0.0]


def compose_geometry(tgt2src, src_k4, tgt_k4, imsize=64):
    """training/utils.py:64-81 — (cat(extrinsics[3x4], src [fx,fy,cx,cy], tgt [fx,fy,cx,cy]) - mean)/std, 0 where std=0."""
    mean = torch.tensor(GEOM_MEAN, dtype=tgt2src.dtype, device=tgt2src.device)
    std = torch.tensor(GEOM_STD, dtype=tgt2src.dtype, device=tgt2src.device)
    mean[12:] *= imsize / 64
    std[12:] *= (imsize / 64) ** 2
    g = torch.cat((tgt2src.reshape(*tgt2src.shape[:-2], 12), src_k4, tgt_k4), -1)
    return torch.where(std > 0, (g - mean) / std, torch.zeros_like(g))


# ----------------------------------------------------------------------------- metric statistics (calculate_metrics.py)


class StatsOracle:
    """fp64 feature statistics of calculate_stats_for_iterable_nvs: update_mu_sigma (:158-172) and reduce (:174-183)."""

    def __init__(self, dim):
        self.cum_mu = np.zeros([dim], dtype=np.float64)
        self.cum_sigma = np.zeros([dim, dim], dtype=np.float64)
        self.n = 0

    def update(self, features, features2=None):
        f = np.asarray(features, dtype=np.float64)
        if features2 is not None:
            f = np.concatenate([f, np.asarray(features2, dtype=np.float64)], axis=-1)
        self.cum_mu += f.sum(0)
        self.cum_sigma += f.T @ f
        self.n += f.shape[0]

    def finalize(self, num_images=None):
        n = self.n if num_images is None else num_images
        mu = self.cum_mu / n
        sigma = (self.cum_sigma - np.outer(mu, mu) * n) / (n - 1)
        return dict(mu=mu, sigma=sigma)


def psnr_u8(images, tgt):
    """calculate_metrics.py:148 — 10 log10(255^2 / mean((x - y)^2)) per image, evaluated in fp64."""
    x = np.asarray(images, dtype=np.float64)
    y = np.asarray(tgt, dtype=np.float64)
    return 10 * np.log10(255.0 ** 2 / ((x - y) ** 2).mean(axis=(1, 2, 3)))


def frechet_distance(mu, sigma, mu_ref, sigma_ref):
    """calculate_metrics.py:309-311."""
    import scipy.linalg
    m = np.square(mu - mu_ref).sum()
    s = scipy.linalg.sqrtm(np.dot(sigma, sigma_ref))     # (scipy >= 1.16 has no `disp`; older ones return the same matrix)
    s = s[0] if isinstance(s, tuple) else s
    return float(np.real(m + np.trace(sigma + sigma_ref - s * 2)))
