"""Plan builder: turns one NVPrecond parameter tree + batch size into a recorded sequence of
libvividb200 ops (include/vivid_b200.h) over fixed device buffers.

Everything in HBM is NHWC in ONE 16-bit format (fp16 by default — the reference's own reduced
precision); accumulation, normalisation statistics and the sampler state are fp32.

Dataflow per block (reference Block.forward, training/models.py:165-206).  `raw` = block output
(the residual stream), `nsilu` = mp_silu(pixel_norm(raw)), `silu(w)` = mp_silu(w*raw):
  enc : conv_res0 reads the PREVIOUS block's nsilu (its pixel-norm + mp_silu were fused into that block's
        last GEMM epilogue); conv_res1's epilogue re-derives pixel_norm(raw_prev) from the TMA-staged
        residual tile, applies mp_sum + clip and emits raw [+ nsilu for the next block] [+ silu(wb) for
        the decoder block that will concatenate this skip].  Levels with C > 256 (one GEMM tile cannot hold
        a whole pixel) use the PIXNORM pass instead; 'down' blocks use the fused pool+norm pass.
  dec : mp_cat is folded into the consumers: conv_res0 and the 1x1 conv_skip run a two-source K loop over
        [silu(wa*x) | silu(wb*skip)] and [x | skip] (mp_cat weights inside mp_silu, resp. folded into the
        conv_skip weights), so no concatenated tensor exists.
  attn: 1x1 qkv GEMM (epilogue: per-head normalise, scatter to [B,h,S,D]) [+ 1x1 kv GEMM on the source-view
        features, written behind the self keys] -> fused attention -> 1x1 proj GEMM (mp_sum, clip).
"""
import ctypes as C
import math

import os

import torch

from . import _lib as L


def _pad(v, m):
    return (v + m - 1) // m * m


class Act:
    """Block output and the derived forms later consumers read (all 16-bit NHWC [B*R*R, C])."""

    def __init__(self, B, R, ch):
        self.B, self.R, self.C = B, R, ch
        self.raw = None       # block output (residual stream / GEMM operand)
        self.norm = None      # pixel_norm(raw)
        self.nsilu = None     # mp_silu(pixel_norm(raw))
        self.rnorm = None     # fp32 per-pixel 1/(eps + rms(raw)), written next to nsilu (viewed as 2 x 16-bit per pixel)
        self.silu = {}        # scale -> mp_silu(scale * raw)
        self.is_skip = False
        self.is_feature = False

    def tensors(self, keep_raw=False):
        out = [] if keep_raw else [self.raw]
        return out + [self.norm, self.nsilu, self.rnorm] + list(self.silu.values())


_DT = {torch.float32: L.VB_F32, torch.float16: L.VB_F16, torch.bfloat16: L.VB_BF16}
AUTOTUNE = os.environ.get("VB_AUTOTUNE", "1") != "0"    # plan-time layout tuning of the conv layers (see Plan._tune_conv)
_TUNE_CACHE = {}
FOLD_RES = os.environ.get("VB_FOLD_RES", "1") != "0"    # mp_sum's coefficient of the GEMM result folded into the prepared weights
FULLROW_MAX = 256          # widest channel count one GEMM tile (and TMEM accumulator buffer) can hold
# K-split (vb_conv_desc.tune bit 8) for the 3x3 convs of the 8x8 level: two CTAs per tile, each over half of the K loop (72-144
# blocks).  It changes the summation order, so the rule below looks at the LAYER only (never the batch size or a timing): results
# stay identical across batch splits and ranks.  Built, parity-tested and measured (profiles/r02_ksplit.txt): the tuner's narrow N
# tiles already fill the chip at batch >= 32 and the hand-over through L2 costs what the shorter loop saves, so it is faster only
# at batch <= 8 and 1.1 % SLOWER per vivid-base call at batch 128 — off unless VB_KSPLIT=1.
KSPLIT = os.environ.get("VB_KSPLIT", "0") == "1"
KSPLIT_MAX_RES = 8


class Plan:
    """A recorded denoiser call for a fixed batch size."""

    def __init__(self, net, B, device):
        self.lib = L.lib()
        L.check(self.lib.vb_device_check(), "vb_device_check")
        self.op_dtype = L.operand_torch_dtype()
        self.op_code = self.lib.vb_operand_dtype()
        self.net = net
        self.device = device
        self.B = B                      # number of target images (outputs)
        self.dual = net.dual
        self.Bx = 2 * B if net.dual else B
        self.keep = []                  # every buffer the plan touches (keeps them alive)
        self.owned_bytes = 0            # device bytes allocated by this plan (prepared weights, activations, I/O)
        self.pool = {}                  # numel -> free activation buffers
        self.op_info = []               # per recorded op: (kind, label, algorithmic flops, algorithmic bytes)
        self.handle = C.c_void_p()
        L.check(self.lib.vb_plan_create(C.byref(self.handle)), "vb_plan_create")
        self.stream = torch.cuda.current_stream(device).cuda_stream
        self.sm_count = torch.cuda.get_device_properties(device).multi_processor_count
        self.alg_flops = 0.0            # algorithmic (unpadded) FLOPs of one call
        self.weight_versions = None
        self.ks_ws = None               # fp32 workspace of the K-split convs
        self._build()

    def __del__(self):
        try:
            if self.handle:
                self.lib.vb_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ buffers
    def buf(self, shape, dtype, zero=False):
        t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
        self.owned_bytes += t.numel() * t.element_size()
        self.keep.append(t)
        return t

    # Activations come from a size-keyed pool: ops replay in order on one stream, so a buffer whose last
    # consumer has been recorded can back a later activation (keeps the working set near L2 / HBM-friendly).
    def act(self, rows, ch):
        free = self.pool.get(rows * ch)
        t = free.pop() if free else self.buf((rows * ch,), self.op_dtype)
        return t.view(rows, ch)

    def to_f32(self, dst, src):
        """dst = float(src): the MPFourier buffers (fp16 in persisted snapshots)."""
        dst.copy_(src.detach().float())

    def fill(self, dst, value):
        dst.fill_(value)

    def release(self, *tensors):
        for t in tensors:
            if t is not None:
                self.pool.setdefault(t.numel(), []).append(t.view(-1))

    def a16(self, B, R, ch):
        return self.act(B * R * R, ch)

    # ------------------------------------------------------------------ weights
    def prep_weight(self, w, gain=1.0, cout_pad=None, perm=(0, 0), split=None, scales=(1.0, 1.0), fp32=False, dst=None):
        """vb_weight_prep: normalise (fp32) + gain + pack; returns the device tensor.  dst (fp32 only): write there instead of
        into a new buffer (the rows of a stacked matrix)."""
        w = w.detach()
        if not w.is_contiguous():
            w = w.contiguous()
        cout, cin = w.shape[0], w.shape[1]
        taps = w.shape[2] * w.shape[3] if w.ndim == 4 else 1
        split = cin if split is None else split
        if fp32:
            if dst is None:
                dst = self.buf((cout, cin * taps), torch.float32)
            assert dst.dtype == torch.float32 and dst.is_contiguous() and dst.numel() == cout * cin * taps
            d = L.WeightPrepDesc(src=w.data_ptr(), dst=dst.data_ptr(), src_dtype=_DT[w.dtype], dst_dtype=L.VB_F32,
                                 cout=cout, cin=cin, taps=taps, cout_pad=cout, split=cin, seg_a_pad=cin, seg_b_pad=0,
                                 perm_parts=0, perm_dim=0, gain=float(gain), scale_a=1.0, scale_b=1.0)
        else:
            sa = _pad(split, 64)
            sb = _pad(cin - split, 64) if cin > split else 0
            cout_pad = cout_pad or _pad(cout, 16)
            dst = self.buf((cout_pad, taps * (sa + sb)), self.op_dtype)
            d = L.WeightPrepDesc(src=w.data_ptr(), dst=dst.data_ptr(), src_dtype=_DT[w.dtype], dst_dtype=self.op_code,
                                 cout=cout, cin=cin, taps=taps, cout_pad=cout_pad, split=split, seg_a_pad=sa,
                                 seg_b_pad=sb, perm_parts=perm[0], perm_dim=perm[1], gain=float(gain),
                                 scale_a=float(scales[0]), scale_b=float(scales[1]))
        self.keep.append(w)
        L.check(self.lib.vb_weight_prep(C.byref(d), self.stream), "vb_weight_prep")
        return dst

    # ------------------------------------------------------------------ op emitters
    def pick_block_n(self, cout_pad, m_pixels, k_blocks, multiple=16, fullrow=False):
        """N tile minimising an estimated makespan: waves x (main loop + epilogue) per tile."""
        if fullrow:
            return cout_pad
        m_tiles = (m_pixels + 127) // 128
        best, best_cost = None, None
        for n in (256, 192, 128, 64, 32, 16):
            if cout_pad % n or n % multiple:
                continue
            tiles = m_tiles * (cout_pad // n)
            waves = -(-tiles // self.sm_count)
            # MMA cycles ~ k_blocks * n/2 (+ fixed issue cost), epilogue ~ 6 cycles per column (overlapped unless it dominates)
            mma = k_blocks * (max(n, 64) / 2.0 + 24)
            epi = 6.0 * n + 400
            cost = waves * max(mma, epi) + min(mma, epi) * 0.15 + 600
            if best_cost is None or cost < best_cost * 0.999:
                best, best_cost = n, cost
        return best

    def _tune_conv(self, d, candidates):
        """Plan-time autotuning of one conv layer: times every legal (block_n, single|pair) candidate on the layer's own
        buffers (batches of launches behind a blocker, best batch) and returns the fastest; the first candidate is the
        heuristic choice and keeps its place unless another is >3 % faster.  The candidates differ in tiling only: every
        output element is computed by the same sequence of MMAs whichever is picked."""
        stream = torch.cuda.current_stream(self.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        for bn, tune in candidates:
            d.block_n, d.tune = bn, tune
            plan = C.c_void_p()
            L.check(self.lib.vb_plan_create(C.byref(plan)), "vb_plan_create")
            try:
                if self.lib.vb_plan_add_conv(plan, C.byref(d)) != 0:
                    continue                       # not a legal layout for this layer (shared-memory budget, tiling)
                for _ in range(2):
                    L.check(self.lib.vb_plan_run(plan, 0, -1, stream.cuda_stream), "vb_plan_run")
                t, reps = float("inf"), 4
                for batch in range(4):
                    L.check(self.lib.vb_spin(80 if batch == 0 else 20 * reps, stream.cuda_stream), "vb_spin")
                    e0.record(stream)
                    for _ in range(reps):
                        L.check(self.lib.vb_plan_run(plan, 0, -1, stream.cuda_stream), "vb_plan_run")
                    e1.record(stream)
                    stream.synchronize()
                    if batch > 0:
                        t = min(t, e0.elapsed_time(e1) / reps)
                    else:           # the first batch only sizes the others: ~0.5 ms of device time each
                        reps = int(min(32, max(4, 0.5 / max(e0.elapsed_time(e1) / reps, 1e-3))))
                if best is None or t < best[0] * 0.97:
                    best = (t, bn, tune)
            finally:
                self.lib.vb_plan_destroy(plan)
        assert best is not None, "no legal conv layout"
        return best[1], best[2]

    def conv(self, x, w, B, R, cin_pad, cout, taps, *, x2=None, cin2_pad=0, cout_pad=None, flags=0, mod=None,
             mod_stride=0, res=None, res_mode=L.VB_RES_NONE, res_t=0.3, clip=None, outs=(), out_f32=None, qkv=None,
             k_real=None, out_rnorm=None, res_rnorm=None, res_folded=False):
        """outs: sequence of (tensor, kind, scale)."""
        cout_pad = cout_pad or _pad(cout, 16)
        fullrow = res_mode == L.VB_RES_PIXNORM or any(k >= L.VB_OUT_NORM for _, k, _ in outs)
        multiple = qkv["D"] if qkv else (64 if outs else 16)
        k_blocks = taps * (cin_pad + cin2_pad) // 64
        bn = self.pick_block_n(cout_pad, B * R * R, k_blocks, multiple, fullrow)
        assert bn is not None and (not fullrow or bn <= FULLROW_MAX), (cout_pad, bn)
        if clip is not None:
            flags |= L.VB_F_CLIP
        if res_folded:
            assert res is not None
            flags |= L.VB_F_RESB_FOLDED
        d = L.ConvDesc(x=x.data_ptr(), x2=L.ptr(x2), w=w.data_ptr(), mod=mod if isinstance(mod, int) else L.ptr(mod),
                       res=L.ptr(res), out_f32=L.ptr(out_f32), out_rnorm=L.ptr(out_rnorm), res_rnorm=L.ptr(res_rnorm), B=B,
                       H=R, W=R, cin_pad=cin_pad, cin2_pad=cin2_pad,
                       cout_pad=cout_pad, taps=taps, block_n=bn, epi_mode=L.VB_EPI_QKVNORM if qkv else L.VB_EPI_PLAIN,
                       flags=flags, mod_stride=mod_stride, ld_f32=cout_pad, res_mode=res_mode, res_t=res_t,
                       clip=clip if clip is not None else 0.0)
        for i, (t, kind, scale) in enumerate(outs):
            d.out[i] = t.data_ptr()
            d.out_kind[i] = kind
            d.out_scale[i] = scale
        if qkv:
            parts = qkv["parts"]
            d.head_dim, d.parts, d.seg_div = qkv["D"], parts, qkv.get("seg_div", 1)
            d.part_ld = qkv.get("ld", 0)
            if parts == 3:      # the softmax's log2(e)/sqrt(D) rides on q: one multiply per q element instead of one per logit
                d.out_scale[0] = math.log2(math.e) / math.sqrt(qkv["D"])
            for j in range(parts):
                d.part_out[j] = qkv["out"][j].data_ptr()
                d.part_seq[j] = qkv["seq"][j]
                d.part_off[j] = qkv["off"][j]
        ks = (KSPLIT and taps == 9 and R <= KSPLIT_MAX_RES and ((cin_pad + cin2_pad) // 64) % 2 == 0 and qkv is None
              and out_f32 is None and len(outs) == 1 and outs[0][1] == L.VB_OUT_RAW and out_rnorm is None
              and ((res_mode == L.VB_RES_NONE and (flags & L.VB_F_MODSILU))
                   or (res_mode == L.VB_RES_PLAIN and not (flags & L.VB_F_MODSILU))))
        if ks:
            need = self.lib.vb_conv_ksplit_ws_bytes(B, R, R, cout_pad)
            if self.ks_ws is None or self.ks_ws.numel() * 4 < need:
                self.ks_ws = self.buf((need // 4,), torch.float32)       # shared by the plan's K-split convs (they replay in order)
            d.ks_ws = self.ks_ws.data_ptr()
            d.tune = 256
        if AUTOTUNE:
            key = (B, R, cin_pad, cin2_pad, cout_pad, taps, flags, res_mode, tuple(k for _, k, _ in outs), out_f32 is not None,
                   (qkv["D"], qkv["parts"], qkv.get("seg_div", 1)) if qkv else None, mod is not None)
            if key not in _TUNE_CACHE:
                ns = [bn] if fullrow else [bn] + [n for n in (256, 192, 128, 64, 32, 16)
                                                  if n != bn and cout_pad % n == 0 and n % multiple == 0]
                # Only bitwise-neutral knobs are tuned: block_n and single/pair change the tiling, not the order in which an
                # output element's K terms are summed, so results stay identical across batch splits and ranks whatever
                # each plan picks (tools/debug_tune_bits.py).  The operand-box sharing mode (tune bits 2-3) and the
                # row-rolling layout (bits 4-5) reorder the K sum (1-ulp flips on ~0.3 % of the elements) and bought
                # nothing measurable, so they stay with the library's deterministic defaults.
                cands = [(n, t) for n in ns for t in (0, 1, 2)]
                # ping-pong epilogue (tune bit 6, block_n <= 128): same arithmetic per element, other schedule
                cands += [(n, t | 64) for n in ns if n <= 128 and outs for t in (0, 1, 2)]
                if ks:          # K-split layers: single-CTA tiles of either width (the split point does not depend on block_n)
                    cands = [(n, 256) for n in ns]
                if taps == 1:
                    # 1x1 layers: the N tile's weights resident per CTA (tune bit 7) — the K order is unchanged as well
                    cands += [(n, t | 128) for n in ns for t in (0, 1, 2)]
                    cands += [(n, t | 64 | 128) for n in ns if n <= 128 and outs for t in (0, 1, 2)]
                _TUNE_CACHE[key] = self._tune_conv(d, cands)
                if os.environ.get("VB_TUNE_LOG"):
                    print("tune", key[:8], "heuristic bn", ns[0], "->", _TUNE_CACHE[key], flush=True)
            bn, d.tune = _TUNE_CACHE[key]
            d.block_n = bn
        L.check(self.lib.vb_plan_add_conv(self.handle, C.byref(d)), "vb_plan_add_conv")
        fl = 2.0 * B * R * R * cout * (k_real if k_real is not None else taps * (cin_pad + cin2_pad))
        self.alg_flops += fl
        P = B * R * R
        by = 2.0 * P * (cin_pad + cin2_pad) + 2.0 * cout_pad * taps * (cin_pad + cin2_pad)
        by += P * cout_pad * (2.0 * (res is not None) + 2.0 * len(outs) + 4.0 * (out_f32 is not None) + (2.0 if qkv else 0.0))
        self.op_info.append(("conv%d" % (3 if taps == 9 else 1),
                             f"{R}x{R} k{taps * (cin_pad + cin2_pad)} n{cout_pad} bn{bn} o{len(outs)}r{res_mode}", fl, by))

    def eltwise(self, kind, a, B, R, ca, *, b=None, cb=0, wa=1.0, wb=1.0, out=None, out_silu=None):
        d = L.EwDesc(a=a.data_ptr(), b=L.ptr(b), out=L.ptr(out), out_silu=L.ptr(out_silu), kind=kind, B=B, H=R, W=R,
                     ca=ca, cb=cb, wa=wa, wb=wb)
        L.check(self.lib.vb_plan_add_eltwise(self.handle, C.byref(d)), "vb_plan_add_eltwise")
        P = B * R * R
        ctot = ca + cb
        rd = 2.0 * P * ctot * (0.25 if kind == L.VB_EW_UP else 4.0 if kind == L.VB_EW_DOWN_PIXNORM else 1.0)
        wr = 2.0 * P * ctot * ((out is not None) + (out_silu is not None))
        self.op_info.append(("eltwise", f"kind{kind} {R}x{R} c{ctot}", 0.0, rd + wr))

    def attention(self, q, k, v, y, B, heads, sq, sk, D, zero_keys, ld=0):
        # q carries log2(e)/sqrt(D) already (folded into the QKV GEMM epilogue, see run_unet): p = 2^(q.k)
        d = L.AttnDesc(q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), y=y.data_ptr(), B=B, heads=heads, sq=sq, sk=sk,
                       head_dim=D, zero_keys=zero_keys, q_prescaled=1, ld=ld)
        L.check(self.lib.vb_plan_add_attn(self.handle, C.byref(d)), "vb_plan_add_attn")
        self.alg_flops += 4.0 * B * heads * sq * sk * D
        self.op_info.append(("attn", f"h{heads} sq{sq} sk{sk} d{D}", 4.0 * B * heads * sq * sk * D,
                             2.0 * B * heads * D * (2 * sq + 2 * sk)))

    # ------------------------------------------------------------------ embedding of one UNet
    def embed(self, unet, B, sigma, sigma_stride, geom, geom_rows, label_dim, noise_scale, geom_scale):
        blocks = [(s, unet.enc[s.name] if s.group == "enc" else unet.dec[s.name])
                  for s in unet.enc_specs + unet.dec_specs if s.kind == "block"]
        offs, total = {}, 0
        for s, _ in blocks:
            offs[(s.group, s.name)] = total
            total += s.cout
        w_mod = self.buf((total, unet.cemb), torch.float32)
        for s, m in blocks:
            o = offs[(s.group, s.name)]
            self.prep_weight(m.emb_linear.weight, gain=float(m.emb_gain.detach().float().item()), fp32=True, dst=w_mod[o:o + s.cout])
        w_noise = self.prep_weight(unet.emb_noise.weight, fp32=True)
        w_label = self.prep_weight(unet.emb_label.weight, fp32=True) if unet.emb_label is not None else None
        freqs = self.buf((unet.cnoise,), torch.float32)
        phases = self.buf((unet.cnoise,), torch.float32)
        self.to_f32(freqs, unet.emb_fourier.freqs)
        self.to_f32(phases, unet.emb_fourier.phases)
        emb = self.buf((B, unet.cemb), torch.float32)
        mod = self.buf((B, total), torch.float32)
        d = L.EmbDesc(sigma=sigma.data_ptr(), geom=L.ptr(geom), freqs=freqs.data_ptr(), phases=phases.data_ptr(),
                      w_noise=w_noise.data_ptr(), w_label=L.ptr(w_label), w_mod=w_mod.data_ptr(), emb=emb.data_ptr(),
                      mod=mod.data_ptr(), B=B, sigma_n=B, sigma_stride=sigma_stride, cnoise=unet.cnoise, cemb=unet.cemb,
                      label_dim=label_dim, mod_total=total, geom_rows=geom_rows, label_balance=unet.label_balance,
                      noise_scale=noise_scale, geom_scale=geom_scale)
        L.check(self.lib.vb_plan_add_embed(self.handle, C.byref(d)), "vb_plan_add_embed")
        self.op_info.append(("embed", f"cemb{unet.cemb} mod{total}", 0.0, 4.0 * total * unet.cemb))
        return mod, offs, total

    # ------------------------------------------------------------------ one UNet / encoder
    @staticmethod
    def sum_coeff(t):
        """mp_sum's coefficient of its second operand (training/models.py:71-72)."""
        return float(t) / math.sqrt((1.0 - float(t)) ** 2 + float(t) ** 2)

    @staticmethod
    def cat_weights(na, nb, t):
        """mp_cat scale factors (training/models.py:78-84)."""
        cc = math.sqrt((na + nb) / ((1 - t) ** 2 + t ** 2))
        return cc / math.sqrt(na) * (1 - t), cc / math.sqrt(nb) * t

    def run_unet(self, unet, x_in, B, mod, offs, mod_total, features=None, feat_seg=1, zero_feature_keys=False,
                 collect_features=False):
        """Emit the ops of UNet.forward.  x_in: 16-bit NHWC [B,R,R,64] (im2col of image + ones channels, zero padded).
        features: list of Act consumed by cross-attention blocks in order.
        Returns (raw network output fp32 [P,16] or None, collected feature Acts)."""
        specs = unet.enc_specs + unet.dec_specs
        t_cat = unet.concat_balance
        # mp_cat weights are needed when the PRODUCERS run (they are folded into mp_silu copies / conv_skip weights)
        pending, skip_scale, cat_scale = [], {}, {}
        width = None
        for s in specs:
            if s.group == "enc":
                pending.append(s)
                width = s.cout
            else:
                if s.skip_ch:
                    e = pending.pop()
                    wa, wb = self.cat_weights(width, e.cout, t_cat)
                    skip_scale[e.name] = wb
                    cat_scale[s.name] = (wa, wb)
                width = s.cout
        feats_out, skips = [], []
        features = list(features or [])
        cur = None

        def want(i):
            """Output forms block i must emit: list of (attribute, kind, scale) in slot order."""
            s = specs[i]
            nxt = specs[i + 1] if i + 1 < len(specs) else None
            fullrow = s.cout <= FULLROW_MAX
            forms = [("raw", L.VB_OUT_RAW, 1.0)]
            if nxt is not None:
                if nxt.flavor == "enc" and nxt.resample == "keep" and not nxt.has_conv_skip and fullrow:
                    forms.append(("nsilu", L.VB_OUT_NORM_SILU, 1.0))
                elif nxt.flavor == "dec" and nxt.resample == "keep" and not nxt.skip_ch:
                    forms.append(("silu", L.VB_OUT_SILU, 1.0))
                elif nxt.flavor == "dec" and nxt.skip_ch:
                    forms.append(("silu", L.VB_OUT_SILU, cat_scale[nxt.name][0]))
            if s.group == "enc" and s.name in skip_scale:
                forms.append(("silu", L.VB_OUT_SILU, skip_scale[s.name]))
            return list(dict.fromkeys(forms))            # e.g. 8x8_block2: silu(1.0) serves both in0 and the first mp_cat

        def alloc_outs(out, forms):
            outs = []
            for attr, kind, scale in forms:
                t = self.a16(out.B, out.R, out.C)
                if attr == "silu":
                    out.silu[scale] = t
                else:
                    setattr(out, attr, t)
                if attr == "nsilu":          # the consumer's residual scale travels with it (VB_RES_SCALED)
                    out.rnorm = self.act(out.B * out.R * out.R, 2)
                outs.append((t, kind, scale))
            return outs

        for i, s in enumerate(specs):
            mod_ = unet.enc[s.name] if s.group == "enc" else unet.dec[s.name]
            out = Act(B, s.res, s.cout)
            out.is_skip = s.group == "enc"
            out.is_feature = collect_features and s.heads > 0
            R, Cc = s.res, s.cout
            temps, popped_skip = [], None
            res_rnorm = None

            if s.kind == "conv":
                # x_in holds the im2col'd 3x3 neighbourhood (vb_precond_in, im2col=1): the first conv is a K=64 1x1 GEMM
                w = self.prep_weight(mod_.weight.detach().reshape(Cc, s.cin * 9, 1, 1))
                outs = alloc_outs(out, want(i))
                self.conv(x_in, w, B, R, 64, Cc, 1, outs=outs, k_real=9 * s.cin, out_rnorm=out.rnorm)
                cur = out
                skips.append(out)
                continue

            fullrow = Cc <= FULLROW_MAX
            x2 = None
            k2 = 0
            # ---------------- main branch: residual base + conv_res0 operand(s)
            if s.flavor == "enc":
                if s.resample == "down":
                    base, a0 = self.a16(B, R, Cc), self.a16(B, R, Cc)
                    temps += [base, a0]
                    self.eltwise(L.VB_EW_DOWN_PIXNORM, cur.raw, B, R, Cc, out=base, out_silu=a0)
                    res, res_mode = base, L.VB_RES_PLAIN
                elif s.has_conv_skip:
                    base, a0 = self.a16(B, R, Cc), self.a16(B, R, Cc)
                    temps += [base, a0]
                    w = self.prep_weight(mod_.conv_skip.weight)
                    if fullrow:      # x = normalize(conv_skip(x)) in one GEMM
                        self.conv(cur.raw, w, B, R, _pad(s.cin, 64), Cc, 1, k_real=s.cin,
                                  outs=[(base, L.VB_OUT_NORM, 1.0), (a0, L.VB_OUT_NORM_SILU, 1.0)])
                    else:
                        tmp = self.a16(B, R, Cc)
                        temps.append(tmp)
                        self.conv(cur.raw, w, B, R, _pad(s.cin, 64), Cc, 1, k_real=s.cin, outs=[(tmp, L.VB_OUT_RAW, 1.0)])
                        self.eltwise(L.VB_EW_PIXNORM, tmp, B, R, Cc, out=base, out_silu=a0)
                    res, res_mode = base, L.VB_RES_PLAIN
                elif cur.nsilu is not None:          # pixel-norm fused on both sides
                    a0, res, res_mode, res_rnorm = cur.nsilu, cur.raw, L.VB_RES_SCALED, cur.rnorm
                else:
                    base, a0 = self.a16(B, R, Cc), self.a16(B, R, Cc)
                    temps += [base, a0]
                    self.eltwise(L.VB_EW_PIXNORM, cur.raw, B, R, Cc, out=base, out_silu=a0)
                    res, res_mode = base, L.VB_RES_PLAIN
                k0 = Cc
            else:
                if s.resample == "up":
                    base, a0 = self.a16(B, R, Cc), self.a16(B, R, Cc)
                    temps += [base, a0]
                    self.eltwise(L.VB_EW_UP, cur.raw, B, R, Cc, out=base, out_silu=a0)
                    res, res_mode, k0 = base, L.VB_RES_PLAIN, Cc
                elif s.skip_ch:
                    skip = skips.pop()
                    popped_skip = skip
                    na, nb = cur.C, skip.C
                    assert nb == s.skip_ch and na + nb == s.cin and na % 64 == 0 and nb % 64 == 0
                    wa, wb = cat_scale[s.name]
                    a0, x2, k0, k2 = cur.silu[wa], skip.silu[wb], na, nb
                    base = self.a16(B, R, Cc)
                    temps.append(base)
                    w = self.prep_weight(mod_.conv_skip.weight, split=na, scales=(wa, wb))
                    self.conv(cur.raw, w, B, R, na, Cc, 1, x2=skip.raw, cin2_pad=nb, outs=[(base, L.VB_OUT_RAW, 1.0)])
                    res, res_mode = base, L.VB_RES_PLAIN
                else:
                    a0, res, res_mode, k0 = cur.silu[1.0], cur.raw, L.VB_RES_PLAIN, Cc
            assert k0 % 64 == 0, f"{s.name}: {k0} input channels are not a multiple of 64"

            # ---------------- residual branch
            y0 = self.a16(B, R, Cc)
            temps.append(y0)
            w0 = self.prep_weight(mod_.conv_res0.weight, split=k0 if x2 is not None else None)
            mo = offs[(s.group, s.name)]
            self.conv(a0, w0, B, R, k0, Cc, 9, x2=x2, cin2_pad=k2, flags=L.VB_F_MODSILU, mod=mod.data_ptr() + 4 * mo,
                      mod_stride=mod_total, outs=[(y0, L.VB_OUT_RAW, 1.0)])
            # mp_sum(x, y, t) = (x (1-t) + y t) / sqrt((1-t)^2 + t^2): y's coefficient rides on the prepared weights of the GEMM that
            # produces y (one fp32 multiply per output element less in the epilogue of every residual layer)
            w1 = self.prep_weight(mod_.conv_res1.weight, gain=self.sum_coeff(mod_.res_balance) if FOLD_RES else 1.0)
            clip = mod_.clip_act
            if s.heads == 0:
                outs = alloc_outs(out, want(i))
                self.conv(y0, w1, B, R, Cc, Cc, 9, res=res, res_mode=res_mode, res_rnorm=res_rnorm, res_t=mod_.res_balance,
                          clip=clip, outs=outs, out_rnorm=out.rnorm, res_folded=FOLD_RES)
            else:
                xr = self.a16(B, R, Cc)
                temps.append(xr)
                self.conv(y0, w1, B, R, Cc, Cc, 9, res=res, res_mode=res_mode, res_rnorm=res_rnorm, res_t=mod_.res_balance,
                          outs=[(xr, L.VB_OUT_RAW, 1.0)], res_folded=FOLD_RES)
                S, D, h = R * R, s.head_dim, s.heads
                nseg = feat_seg if s.xattn else 0
                real_seg = 0 if zero_feature_keys else nseg
                sk = S * (1 + real_seg)
                # D = 32 (the SR UNet): rows zero-padded to 64 elements so that the tcgen05 attention kernel (64-wide
                # operand rows) serves them; the buffers are private to this layer and zeroed once — the GEMM epilogue only
                # ever writes the lower 32 elements.  (Twice the attention FLOPs of a dense D = 32 kernel, 0.7 % of the SR net.)
                ld = 64 if (D == 32 and S % 256 == 0 and sk % 128 == 0) else 0
                if ld:
                    q, k, v = (self.buf((B * h * n, ld), self.op_dtype, zero=True) for n in (S, sk, sk))
                else:
                    q = self.act(B * h * S, D)
                    k = self.act(B * h * sk, D)
                    v = self.act(B * h * sk, D)
                wq = self.prep_weight(mod_.attn_qkv.weight, perm=(3, D))
                self.conv(xr, wq, B, R, Cc, 3 * Cc, 1,
                          qkv=dict(D=D, parts=3, out=[q, k, v], seq=[S, sk, sk], off=[0, 0, 0], ld=ld))
                if s.xattn and not zero_feature_keys:
                    f = features.pop(0)
                    assert f.C == Cc and f.R == R, f"{s.name}: feature map mismatch"
                    wkv = self.prep_weight(mod_.x_attn_kv.weight, perm=(2, D))
                    self.conv(f.raw, wkv, f.B, R, Cc, 2 * Cc, 1,
                              qkv=dict(D=D, parts=2, out=[k, v], seq=[sk, sk], off=[S, S], seg_div=feat_seg, ld=ld))
                    # f.raw stays allocated for the life of the plan: return_features / inject_features read and write it
                y = self.a16(B, R, Cc)
                temps += [y] if ld else [q, k, v, y]           # (padded q/k/v are private: never recycled through the pool)
                # unconditional model: x_attn_kv(0) == 0 -> the S*nseg zero keys are accounted for analytically
                self.attention(q, k, v, y, B, h, S, sk, D, S * nseg if zero_feature_keys else 0, ld=ld)
                wp = self.prep_weight(mod_.attn_proj.weight, gain=self.sum_coeff(mod_.attn_balance) if FOLD_RES else 1.0)
                outs = alloc_outs(out, want(i))
                self.conv(y, wp, B, R, Cc, Cc, 1, res=xr, res_mode=L.VB_RES_PLAIN, res_t=mod_.attn_balance, clip=clip,
                          outs=outs, out_rnorm=out.rnorm, res_folded=FOLD_RES)
            if out.is_feature:
                feats_out.append(out)
            if s.group == "enc":
                skips.append(out)
            # recycle: this block's temporaries, the consumed skip, and the previous block's output unless it
            # lives on as a skip connection (encoder outputs) or as a source-view feature
            self.release(*temps)
            if popped_skip is not None:
                self.release(*popped_skip.tensors(keep_raw=popped_skip.is_feature))
            if cur is not None and not cur.is_skip:
                self.release(*cur.tensors(keep_raw=cur.is_feature))
            cur = out

        raw = None
        if unet.out_conv is not None:
            wo = self.prep_weight(unet.out_conv.weight, gain=float(unet.out_gain.detach().float().item()), cout_pad=16)
            raw = self.buf((B * cur.R * cur.R, 16), torch.float32)
            self.conv(cur.raw, wo, B, cur.R, cur.C, unet.out_conv.out_channels, 9, cout_pad=16, out_f32=raw)
        return raw, feats_out

    # ------------------------------------------------------------------ whole NVPrecond call
    def _build(self):
        net, B, Bx = self.net, self.B, self.Bx
        R = net.img_resolution
        sd = float(net.sigma_data)
        # persistent I/O buffers (callers copy into / out of these)
        self.in_x = self.buf((Bx, 3, R, R), torch.float32, zero=True)
        self.in_src = self.buf((Bx, 3, R, R), torch.float32, zero=True) if net.encoder is not None else None
        self.in_sigma = self.buf((Bx,), torch.float32)
        self.fill(self.in_sigma, 1.0)
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
            src16 = self.buf((Bx * R * R, 64), self.op_dtype)
            d = L.PrecondInDesc(x=self.in_src.data_ptr(), cond=None, noise=None, sigma=None, out=src16.data_ptr(), B=Bx,
                                R=R, cpad=64, sigma_n=1, sigma_stride=0, im2col=1, img_stride=3 * R * R, sigma_data=sd, noisy_sr=0.0)
            L.check(self.lib.vb_plan_add_precond_in(self.handle, C.byref(d)), "vb_plan_add_precond_in")
            self.op_info.append(("precond", "in", 0.0, d.B * R * R * (12.0 + 128.0)))
            mod, offs, total = self.embed(enc, Bx, self.in_sigma, 1, self.in_geom if ldim_enc else None, Bx, ldim_enc,
                                          0.0 if net.no_time_enc else 1.0, geom_scale)
            _, features = self.run_unet(enc, src16, Bx, mod, offs, total, collect_features=True)
            feat_seg = 2 if self.dual else 1
        # ops [0, enc_ops) are the source-view encoder; its outputs (self.features) are what the reference's
        # return_features / inject_features hand around (training/models.py:664-672, snapshot :612-626)
        self.enc_ops = self.lib.vb_plan_num_ops(self.handle)
        self.features = list(features or [])

        unet = net.unet
        x16 = self.buf((B * R * R, 64), self.op_dtype)
        step = 2 if self.dual else 1
        d = L.PrecondInDesc(x=self.in_x.data_ptr(), cond=L.ptr(self.in_cond), noise=L.ptr(self.in_noise),
                            sigma=self.in_sigma.data_ptr(), out=x16.data_ptr(), B=B, R=R, cpad=64, sigma_n=B,
                            sigma_stride=step, im2col=1, img_stride=3 * R * R * step, sigma_data=sd,
                            noisy_sr=float(net.noisy_sr if net.noisy_sr is not None else 0.0))
        L.check(self.lib.vb_plan_add_precond_in(self.handle, C.byref(d)), "vb_plan_add_precond_in")
        self.op_info.append(("precond", "in", 0.0, d.B * R * R * (12.0 + 128.0)))
        mod, offs, total = self.embed(unet, B, self.in_sigma, step, self.in_geom if ldim_unet else None, B, ldim_unet,
                                      1.0, geom_scale)
        raw, _ = self.run_unet(unet, x16, B, mod, offs, total, features=features, feat_seg=feat_seg,
                               zero_feature_keys=net.encoder is None)
        d = L.PrecondOutDesc(x=self.in_x.data_ptr(), f=raw.data_ptr(), sigma=self.in_sigma.data_ptr(),
                             d_out=self.out_d.data_ptr(), B=B, R=R, ldf=16, sigma_n=B, sigma_stride=step,
                             img_stride=3 * R * R * step, sigma_data=sd)
        L.check(self.lib.vb_plan_add_precond_out(self.handle, C.byref(d)), "vb_plan_add_precond_out")
        self.op_info.append(("precond", "out", 0.0, B * R * R * (12.0 + 64.0 + 12.0)))
        self.num_ops = self.lib.vb_plan_num_ops(self.handle)
        # whole-call C entry point (vb_denoise): any host can now run this plan with plain device pointers
        io = L.IoDesc(in_x=self.in_x.data_ptr(), in_src=L.ptr(self.in_src), in_sigma=self.in_sigma.data_ptr(),
                      in_geom=self.in_geom.data_ptr(), in_cond=L.ptr(self.in_cond), in_noise=L.ptr(self.in_noise),
                      out_d=self.out_d.data_ptr(), n_x=Bx, n_out=B, img_elems=3 * R * R, geom_dim=self.in_geom.shape[1],
                      workspace_bytes=self.owned_bytes)
        L.check(self.lib.vb_plan_bind_io(self.handle, C.byref(io)), "vb_plan_bind_io")
        self.launches = int(self.lib.vb_plan_query(self.handle, 1))
        self.padded_flops = self.lib.vb_plan_query(self.handle, 0)

    # ------------------------------------------------------------------ execution
    def profile(self, repeats=3):
        """Time every recorded op on its own with CUDA events (eager replay, launching stream = current stream).
        Returns [(kind, label, flops, bytes, milliseconds)] — the source of the roofline numbers in bench.py."""
        assert len(self.op_info) == self.num_ops, (len(self.op_info), self.num_ops)
        stream = torch.cuda.current_stream(self.device)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(self.num_ops + 1)]
        best = [float("inf")] * self.num_ops
        for _ in range(repeats):
            # a blocker holds the stream while the host enqueues the whole plan: the ops then run back to back and the
            # events see device time only (launched one by one, every op under ~12 us measured the host's launch rate)
            L.check(self.lib.vb_spin(max(2000, 40 * self.num_ops), stream.cuda_stream), "vb_spin")
            ev[0].record(stream)
            for i in range(self.num_ops):
                L.check(self.lib.vb_plan_run(self.handle, i, i + 1, stream.cuda_stream), "vb_plan_run")
                ev[i + 1].record(stream)
            stream.synchronize()
            for i in range(self.num_ops):
                best[i] = min(best[i], ev[i].elapsed_time(ev[i + 1]))
        return [(k, lab, fl, by, ms) for (k, lab, fl, by), ms in zip(self.op_info, best)]

    def run(self, graph=True, section="all"):
        """Replay the plan: everything, the source-view encoder alone ('enc') or the denoising UNet alone ('unet')."""
        first, last = {"all": (0, -1), "enc": (0, self.enc_ops), "unet": (self.enc_ops, -1)}[section]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if graph:
            L.check(self.lib.vb_plan_launch_graph_range(self.handle, first, last, stream), "vb_plan_launch_graph_range")
        else:
            L.check(self.lib.vb_plan_run(self.handle, first, last, stream), "vb_plan_run")

    def feature_views(self):
        """The encoder's cross-attention feature maps as logical NCHW [Bx, C, R, R] views of the 16-bit NHWC buffers."""
        return [f.raw.view(f.B, f.R, f.R, -1)[..., :f.C].permute(0, 3, 1, 2) for f in self.features]
