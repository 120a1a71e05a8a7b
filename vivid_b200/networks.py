"""Host-side mirror of the reference's network objects for the guided denoising path.

The classes here keep the reference's names, constructor arguments, attribute names and
state_dict keys (training/models.py:107-126 MPConv, :89-101 MPFourier, :131-206 Block,
:211-315 XAttnBlock, :320-406 UNet, :411-518 XAttnUNet, :523-570 UNetEncoder, :575-582
SRXAttnUNet, :589-749 NVPrecond; snapshot experiments/code/training/models.py for the vanilla
semantics) so a reference `state_dict()` loads unchanged and `generate_images_nvs(net=...,
gnet=..., sr_model=...)` can be handed these modules.  They are PARAMETER CONTAINERS: the
arithmetic runs in libvividb200.so through the plans built by `vivid_b200.engine`.

The layer table is produced declaratively by `unet_layout()` (one BlockSpec per entry of the
reference's enc/dec ModuleDicts) instead of the reference's imperative constructors.
"""
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch


@dataclass
class BlockSpec:
    name: str                 # e.g. '32x32_block1'
    group: str                # 'enc' | 'dec'
    kind: str                 # 'conv' (first MPConv) | 'block'
    res: int                  # OUTPUT resolution
    cin: int
    cout: int
    flavor: str = "enc"
    resample: str = "keep"    # 'keep' | 'up' | 'down'
    heads: int = 0
    head_dim: int = 0
    xattn: bool = False       # consumes source-view features (XAttnBlock)
    skip_ch: int = 0          # channels of the skip concatenated in front of a decoder block (0 = none)

    @property
    def has_conv_skip(self):
        return self.kind == "block" and self.cin != self.cout


def unet_layout(img_resolution, in_channels, model_channels=192, channel_mult=(1, 2, 3, 4), num_blocks=3,
                attn_resolutions=(16, 8), extra_attn=None, channels_per_head=64, xattn=False):
    """Layer table of UNet / XAttnUNet (reference training/models.py:340-383 / :438-480)."""
    widths = [model_channels * m for m in channel_mult]
    top = len(widths) - 1

    def attention(res, slot, level):
        return res in attn_resolutions or (extra_attn is not None and extra_attn == slot and level != 0)

    def mk(name, group, cin, cout, res, flavor, resample="keep", attn=False):
        heads = cout // channels_per_head if attn else 0
        if attn and heads == 0:
            raise ValueError(f"{name}: {cout} channels cannot host a {channels_per_head}-wide attention head")
        return BlockSpec(name, group, "block", res, cin, cout, flavor, resample, heads,
                         cout // heads if heads else 0, xattn and heads > 0)

    enc: List[BlockSpec] = []
    width = in_channels
    for level, ch in enumerate(widths):
        res = img_resolution >> level
        if level == 0:
            enc.append(BlockSpec(f"{res}x{res}_conv", "enc", "conv", res, width, ch))
            width = ch
        else:
            enc.append(mk(f"{res}x{res}_down", "enc", width, width, res, "enc", "down"))
        for idx in range(num_blocks):
            enc.append(mk(f"{res}x{res}_block{idx}", "enc", width, ch, res, "enc", attn=attention(res, idx, level)))
            width = ch

    dec: List[BlockSpec] = []
    pending = [b.cout for b in enc]
    for level in range(top, -1, -1):
        ch, res = widths[level], img_resolution >> level
        if level == top:
            dec.append(mk(f"{res}x{res}_in0", "dec", width, width, res, "dec", attn=True))
            dec.append(mk(f"{res}x{res}_in1", "dec", width, width, res, "dec"))
        else:
            dec.append(mk(f"{res}x{res}_up", "dec", width, width, res, "dec", "up"))
        for idx in range(num_blocks + 1):
            skip = pending.pop()
            b = mk(f"{res}x{res}_block{idx}", "dec", width + skip, ch, res, "dec",
                   attn=attention(res, num_blocks - idx, level))
            b.skip_ch = skip
            dec.append(b)
            width = ch
    return enc, dec, widths


class MPFourier(torch.nn.Module):
    """Buffers of the magnitude-preserving Fourier features (reference :89-101)."""

    def __init__(self, num_channels, bandwidth=1):
        super().__init__()
        self.register_buffer("freqs", 2 * np.pi * torch.randn(num_channels) * bandwidth)
        self.register_buffer("phases", 2 * np.pi * torch.rand(num_channels))


class MPConv(torch.nn.Module):
    """Weight holder of a magnitude-preserving conv / linear layer (reference :107-126); no bias."""

    def __init__(self, in_channels, out_channels, kernel):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.weight = torch.nn.Parameter(torch.randn(out_channels, in_channels, *kernel))


class Block(torch.nn.Module):
    """Parameters of Block / XAttnBlock (reference :131-206 / :211-315)."""

    def __init__(self, spec: BlockSpec, emb_channels, dropout=0, res_balance=0.3, attn_balance=0.3, clip_act=256):
        super().__init__()
        self.spec = spec
        self.out_channels = spec.cout
        self.flavor = spec.flavor
        self.resample_mode = spec.resample
        self.num_heads = spec.heads
        self.dropout = dropout
        self.res_balance = res_balance
        self.attn_balance = attn_balance
        self.clip_act = clip_act
        self.emb_gain = torch.nn.Parameter(torch.zeros([]))
        self.conv_res0 = MPConv(spec.cout if spec.flavor == "enc" else spec.cin, spec.cout, kernel=[3, 3])
        self.emb_linear = MPConv(emb_channels, spec.cout, kernel=[])
        self.conv_res1 = MPConv(spec.cout, spec.cout, kernel=[3, 3])
        self.conv_skip = MPConv(spec.cin, spec.cout, kernel=[1, 1]) if spec.cin != spec.cout else None
        self.attn_qkv = MPConv(spec.cout, spec.cout * 3, kernel=[1, 1]) if spec.heads else None
        if spec.xattn:
            self.x_attn_kv = MPConv(spec.cout, spec.cout * 2, kernel=[1, 1])
        self.attn_proj = MPConv(spec.cout, spec.cout, kernel=[1, 1]) if spec.heads else None


XAttnBlock = Block


class UNet(torch.nn.Module):
    """Parameter tree of UNet / XAttnUNet / SRXAttnUNet / UNetEncoder."""

    def __init__(self, img_resolution, img_channels, label_dim, model_channels=192, channel_mult=(1, 2, 3, 4),
                 channel_mult_noise=None, channel_mult_emb=None, num_blocks=3, attn_resolutions=(16, 8),
                 label_balance=0.5, concat_balance=0.5, extra_attn=None, epipolar_attention_bias=False,
                 channels_per_head=64, xattn=False, extra_in_channels=0, out_channels=None, encoder_only=False,
                 resample_filter=(1, 1), **block_kwargs):
        super().__init__()
        if epipolar_attention_bias:
            raise NotImplementedError("epipolar_attention_bias is outside the B200 hot path (all presets disable it)")
        if list(resample_filter) != [1, 1]:
            raise NotImplementedError("only resample_filter=[1,1] (2x2 mean pool / nearest) is implemented")
        enc, dec, widths = unet_layout(img_resolution, img_channels + 1 + extra_in_channels, model_channels,
                                       list(channel_mult), num_blocks, list(attn_resolutions), extra_attn,
                                       channels_per_head, xattn)
        self.img_resolution = img_resolution
        self.label_dim = label_dim
        self.model_channels, self.channel_mult, self.num_blocks = model_channels, list(channel_mult), num_blocks
        self.attn_resolutions, self.extra_attn = list(attn_resolutions), extra_attn
        self.cnoise = model_channels * channel_mult_noise if channel_mult_noise is not None else widths[0]
        self.cemb = model_channels * channel_mult_emb if channel_mult_emb is not None else max(widths)
        self.label_balance = label_balance
        self.concat_balance = concat_balance
        self.channels_per_head = channels_per_head
        self.emb_fourier = MPFourier(self.cnoise)
        self.emb_noise = MPConv(self.cnoise, self.cemb, kernel=[])
        self.emb_label = MPConv(label_dim, self.cemb, kernel=[]) if label_dim != 0 else None
        if encoder_only:
            # UNetEncoder (reference :523-534): trailing decoder blocks without attention are dropped
            while dec and dec[-1].heads == 0:
                dec.pop()
        self.enc_specs, self.dec_specs = enc, dec
        self.enc = torch.nn.ModuleDict()
        for s in enc:
            self.enc[s.name] = MPConv(s.cin, s.cout, kernel=[3, 3]) if s.kind == "conv" else Block(s, self.cemb, **block_kwargs)
        self.dec = torch.nn.ModuleDict()
        for s in dec:
            self.dec[s.name] = Block(s, self.cemb, **block_kwargs)
        if not encoder_only:
            self.out_gain = torch.nn.Parameter(torch.zeros([]))
            self.out_conv = MPConv(dec[-1].cout, out_channels if out_channels is not None else img_channels, kernel=[3, 3])
        else:
            self.out_gain = None
            self.out_conv = None

    def feature_specs(self):
        """Specs of the blocks whose outputs (encoder) / inputs (unet) are the source-view features."""
        return [s for s in self.enc_specs + self.dec_specs if s.heads > 0]


def XAttnUNet(img_resolution, img_channels, label_dim, **kw):
    return UNet(img_resolution, img_channels, label_dim, xattn=True, out_channels=3, **kw)


def SRXAttnUNet(img_resolution, img_channels, label_dim, **kw):
    # reference :575-582 — 32-wide heads; first conv also sees the noised low-res image: 2*3+1 channels
    kw.pop("channels_per_head", None)
    return UNet(img_resolution, img_channels, label_dim, xattn=True, out_channels=3, channels_per_head=32,
                extra_in_channels=img_channels, **kw)


def UNetEncoder(img_resolution, img_channels, label_dim, **kw):
    kw.pop("no_cam", None)
    return UNet(img_resolution, img_channels, label_dim, encoder_only=True, **kw)
