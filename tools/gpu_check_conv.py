"""GPU bring-up check for the tcgen05 implicit-GEMM conv kernel + weight prep (run under gpurun).

Compares vb_conv against torch conv2d evaluated in fp32 on the SAME bf16-rounded operands,
so the only difference left is accumulation order (expected rel-L2 ~1e-6).
"""
import ctypes as C
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vivid_b200 import _lib as L  # noqa: E402

h = C.CDLL(L.LIB_PATH)
for name in ("vb_last_error", "vb_weight_prep", "vb_conv", "vb_device_check"):
    res, args = L.SIGNATURES[name]
    getattr(h, name).restype = res
    getattr(h, name).argtypes = args


def chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: rc={rc} {h.vb_last_error().decode()}")


dev = torch.device("cuda")
chk(h.vb_device_check(), "device_check")
stream = torch.cuda.current_stream().cuda_stream


def pad_to(v, m):
    return (v + m - 1) // m * m


def prep_weight(w, gain=1.0, cout_pad=None, perm=(0, 0), split=None, scales=(1.0, 1.0)):
    cout, cin = w.shape[:2]
    taps = w[0, 0].numel() if w.ndim == 4 else 1
    split = cin if split is None else split
    sa, sb = pad_to(split, 64), pad_to(cin - split, 64) if cin > split else 0
    cout_pad = cout_pad or pad_to(cout, 16)
    dst = torch.empty(cout_pad, taps, sa + sb, dtype=torch.bfloat16, device=dev)
    d = L.WeightPrepDesc(src=w.data_ptr(), dst=dst.data_ptr(), src_dtype={torch.float32: 0, torch.float16: 1}[w.dtype],
                         dst_dtype=L.VB_BF16, cout=cout, cin=cin, taps=taps, cout_pad=cout_pad, split=split,
                         seg_a_pad=sa, seg_b_pad=sb, perm_parts=perm[0], perm_dim=perm[1], gain=gain,
                         scale_a=scales[0], scale_b=scales[1])
    chk(h.vb_weight_prep(C.byref(d), stream), "weight_prep")
    return dst


def ref_weight(w, gain=1.0):
    w32 = w.float()
    K = w32[0].numel()
    n = w32.flatten(1).norm(dim=1).reshape(-1, *([1] * (w.ndim - 1)))
    return gain * w32 / (1e-4 * math.sqrt(K) + n)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


fails = 0


def report(name, err, tol):
    global fails
    ok = err <= tol and err == err
    print(f"{'PASS' if ok else 'FAIL'} {name}: rel_l2={err:.3e} (tol {tol:.0e})", flush=True)
    if not ok:
        fails += 1


def run_plain(B, R, cin, cout, taps, block_n, flags=0, gain=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    cin_pad = pad_to(cin, 64)
    cout_pad = pad_to(cout, block_n)
    k = 3 if taps == 9 else 1
    x = torch.randn(B, cin, R, R, generator=g).to(dev)
    w = torch.randn(cout, cin, k, k, generator=g).to(dev)
    x_nhwc = torch.zeros(B, R, R, cin_pad, dtype=torch.bfloat16, device=dev)
    x_nhwc[..., :cin] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    wp = prep_weight(w, gain=gain, cout_pad=cout_pad)
    # reference on identical bf16-rounded operands
    xr = x_nhwc[..., :cin].float().permute(0, 3, 1, 2)
    wr = ref_weight(w, gain).to(torch.bfloat16).float()
    # check the prepared weight itself
    wp_ref = torch.zeros_like(wp, dtype=torch.float32)
    wp_ref[:cout, :, :cin] = wr.reshape(cout, cin, taps).permute(0, 2, 1)
    werr = (wp.float() - wp_ref).abs().max().item()
    y = torch.nn.functional.conv2d(xr, wr, padding=k // 2)
    mod = res = None
    if flags & L.VB_F_MODSILU:
        mod = (torch.randn(B, cout_pad, generator=g) * 0.3 + 1).to(dev)
        y = torch.nn.functional.silu(y * mod[:, :cout, None, None]) / 0.596
    if flags & L.VB_F_RESIDUAL:
        res = torch.randn(B, R, R, cout_pad, generator=g).to(dev)
        t = 0.3
        y = (res[..., :cout].permute(0, 3, 1, 2) * (1 - t) + y * t) / math.sqrt((1 - t) ** 2 + t ** 2)
    if flags & L.VB_F_CLIP:
        y = y.clamp(-1.5, 1.5)
    out32 = torch.full((B, R, R, cout_pad), float("nan"), device=dev)
    out16 = torch.zeros(B, R, R, cout_pad, dtype=torch.bfloat16, device=dev)
    outs = torch.zeros(B, R, R, cout_pad, dtype=torch.bfloat16, device=dev)
    d = L.ConvDesc(x=x_nhwc.data_ptr(), w=wp.data_ptr(), mod=L.ptr(mod), res=L.ptr(res), out_f32=out32.data_ptr(),
                   out_bf16=out16.data_ptr(), out_silu=outs.data_ptr(), B=B, H=R, W=R, cin_pad=cin_pad, cin2_pad=0,
                   cout_pad=cout_pad, taps=taps, block_n=block_n, epi_mode=L.VB_EPI_PLAIN, flags=flags,
                   mod_stride=cout_pad, ld_res=cout_pad, ld_f32=cout_pad, ld_bf16=cout_pad, ld_silu=cout_pad,
                   res_t=0.3, clip=1.5)
    chk(h.vb_conv(C.byref(d), stream), "conv")
    torch.cuda.synchronize()
    got = out32[..., :cout].permute(0, 3, 1, 2)
    name = f"conv B{B} R{R} cin{cin} cout{cout} taps{taps} bn{block_n} flags{flags}"
    report(name + " [w max abs %.1e]" % werr, rel(got, y), 2e-5)
    report(name + " bf16", rel(out16[..., :cout].permute(0, 3, 1, 2).float(), y), 6e-3)
    report(name + " silu", rel(outs[..., :cout].permute(0, 3, 1, 2).float(), torch.nn.functional.silu(y) / 0.596), 8e-3)
    if cout_pad > cout:
        z = out32[..., cout:].abs().max().item()
        report(name + " pad-zero", z, 0.0)


def run_qkv(B, R, C_, heads, D, parts, seg_div=1, block_n=128):
    g = torch.Generator(device="cpu").manual_seed(1)
    cout = heads * parts * D
    x = torch.randn(B, C_, R, R, generator=g).to(dev)
    w = torch.randn(cout, C_, 1, 1, generator=g).to(dev)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    wp = prep_weight(w, cout_pad=cout, perm=(parts, D))
    S = R * R
    Bo = B // seg_div
    seq = [S if j == 0 and parts == 3 else S * (1 + seg_div) for j in range(parts)]
    off = [0 if parts == 3 else S for j in range(parts)]
    outs = [torch.zeros(Bo, heads, seq[j], D, dtype=torch.bfloat16, device=dev) for j in range(parts)]
    po = (C.c_void_p * 3)(*[o.data_ptr() for o in outs] + [None] * (3 - parts))
    d = L.ConvDesc(x=x_nhwc.data_ptr(), w=wp.data_ptr(), part_out=po, B=B, H=R, W=R, cin_pad=C_, cin2_pad=0,
                   cout_pad=cout, taps=1, block_n=block_n, epi_mode=L.VB_EPI_QKVNORM, flags=0, head_dim=D,
                   parts=parts, seg_div=seg_div, part_seq=(C.c_int32 * 3)(*(seq + [0] * (3 - parts))),
                   part_off=(C.c_int32 * 3)(*(off + [0] * (3 - parts))))
    chk(h.vb_conv(C.byref(d), stream), "conv qkv")
    torch.cuda.synchronize()
    xr = x_nhwc.float().permute(0, 3, 1, 2)
    wr = ref_weight(w).to(torch.bfloat16).float()
    y = torch.nn.functional.conv2d(xr, wr)                     # [B, cout, R, R], channel = h*P*D + d*P + j
    y = y.reshape(B, heads, D, parts, S)
    nrm = y.norm(dim=2, keepdim=True)
    y = y / (1e-4 + nrm / math.sqrt(D))
    for j in range(parts):
        ref = y[:, :, :, j, :].permute(0, 1, 3, 2)            # [B, h, S, D]
        if seg_div == 1:
            got = outs[j][:, :, off[j]:off[j] + S].float()
            report(f"qkv B{B} R{R} C{C_} h{heads} D{D} P{parts} part{j}", rel(got, ref), 6e-3)
        else:
            for sgi in range(seg_div):
                got = outs[j][:, :, off[j] + sgi * S: off[j] + (sgi + 1) * S].float()
                report(f"kv-dual seg{sgi} part{j}", rel(got, ref[sgi::seg_div]), 6e-3)


t0 = time.time()
# smallest first: one tile, one k-block per tap
run_plain(1, 16, 64, 64, 1, 64)
run_plain(1, 16, 64, 64, 9, 64)
run_plain(2, 16, 128, 128, 9, 128)
run_plain(2, 32, 64, 128, 9, 128, flags=1)
run_plain(2, 64, 128, 128, 9, 128, flags=2)
run_plain(3, 8, 192, 256, 9, 256, flags=7)          # odd batch with bn=2
run_plain(5, 4, 64, 64, 9, 64, flags=6)             # bn=8 tile, partial
run_plain(2, 16, 320, 192, 1, 192)                  # N=192, multi k-chunk 1x1
run_plain(2, 64, 4, 128, 9, 128)                    # input conv (cin padded 4->64)
run_plain(2, 64, 128, 3, 9, 16, gain=0.7)           # out_conv (cout padded 3->16)
run_plain(1, 128, 64, 64, 9, 64)                    # 128-wide rows
run_plain(1, 256, 64, 64, 9, 64, flags=7)           # SR resolution
run_plain(40, 16, 384, 384, 9, 128, flags=7)        # many tiles per CTA? (40*2*3=240 tiles)
run_plain(64, 32, 256, 256, 9, 128, flags=7)        # 512*2 tiles -> persistent loop, ring wrap
run_qkv(2, 16, 128, 2, 64, 3)
run_qkv(2, 8, 256, 4, 64, 2)
run_qkv(4, 8, 256, 4, 64, 2, seg_div=2)
run_qkv(2, 32, 256, 8, 32, 3, block_n=64)
print(f"done in {time.time() - t0:.1f}s, fails={fails}", flush=True)

# quick timing of a big layer (informational)
if fails == 0:
    B, R, cin, cout = 32, 64, 128, 128
    x = torch.randn(B, R, R, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device=dev)
    wp = prep_weight(w)
    out16 = torch.zeros(B, R, R, cout, dtype=torch.bfloat16, device=dev)
    for bn_ in (64, 128):
        d = L.ConvDesc(x=x.data_ptr(), w=wp.data_ptr(), out_bf16=out16.data_ptr(), B=B, H=R, W=R, cin_pad=cin,
                       cin2_pad=0, cout_pad=cout, taps=9, block_n=bn_, epi_mode=0, flags=0, ld_bf16=cout)
        for _ in range(3):
            chk(h.vb_conv(C.byref(d), stream), "conv")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            chk(h.vb_conv(C.byref(d), stream), "conv")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * B * R * R * cout * cin * 9
        print(f"timing 64x64x128->128 3x3 B32 bn{bn_}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s (incl. tmap encode+launch)")
    B, R, cin, cout = 32, 32, 256, 256
    x = torch.randn(B, R, R, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device=dev)
    wp = prep_weight(w)
    out16 = torch.zeros(B, R, R, cout, dtype=torch.bfloat16, device=dev)
    for bn_ in (128, 256):
        d = L.ConvDesc(x=x.data_ptr(), w=wp.data_ptr(), out_bf16=out16.data_ptr(), B=B, H=R, W=R, cin_pad=cin,
                       cin2_pad=0, cout_pad=cout, taps=9, block_n=bn_, epi_mode=0, flags=0, ld_bf16=cout)
        for _ in range(3):
            chk(h.vb_conv(C.byref(d), stream), "conv")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            chk(h.vb_conv(C.byref(d), stream), "conv")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * B * R * R * cout * cin * 9
        print(f"timing 32x32x256->256 3x3 B32 bn{bn_}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s")
sys.exit(1 if fails else 0)
