"""Micro-benchmark of vb_conv on the layer shapes that dominate the step (B=32 unless given).

Env: VB_B batch, VB_REPS repetitions, VB_ONLY comma list of shape indices, VB_EPI comma list of epilogues:
  simple  one raw output                         mod    conv_res0: modulation + mp_silu -> raw
  r1s     mp_sum(res) + clip -> raw, silu        r2ns   mp_sum(pixel_norm(res)) + clip -> raw, norm_silu
  r3ns / r3nss  as r2ns / r2nss with the residual scaled per pixel (VB_RES_SCALED, the plans' form), b folded into the weights
  r2nss   ... -> raw, norm_silu, silu            (library knobs: VB_TAP_MODE, VB_DBG, VB_GENERIC_EPI)
  qkv     1x1 only: per-head normalise + scatter to [B,h,S,64] (cout = heads*3*64)          VB_TUNE: vb_conv_desc.tune
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vivid_b200 import _lib as L  # noqa: E402

lib = L.lib()
dev = torch.device("cuda")
stream = torch.cuda.current_stream().cuda_stream
SHAPES = [  # R, cin, cout, taps, bn
    (256, 64, 64, 9, 64), (128, 128, 128, 9, 128), (64, 128, 128, 9, 128), (32, 256, 256, 9, 256),
    (16, 384, 384, 9, 128), (8, 512, 512, 9, 64), (8, 512, 512, 9, 128), (16, 384, 1152, 1, 192), (256, 128, 64, 1, 64),
    (256, 128, 64, 9, 64), (64, 192, 192, 9, 192), (16, 384, 384, 9, 192), (64, 128, 128, 1, 128),
    (8, 384, 512, 1, 64), (16, 384, 384, 1, 192), (8, 512, 512, 1, 64),
    (8, 512, 1536, 1, 192), (32, 256, 768, 1, 256), (16, 384, 768, 1, 192), (32, 256, 256, 1, 256), (8, 512, 512, 1, 128),
]
B = int(os.environ.get("VB_B", "32"))
reps = int(os.environ.get("VB_REPS", "10"))
only = os.environ.get("VB_ONLY")
epis = os.environ.get("VB_EPI", "r2ns").split(",")
for idx, (R, cin, cout, taps, bn) in enumerate(SHAPES):
    if only is not None and str(idx) not in only.split(","):
        continue
    dt = L.operand_torch_dtype()
    x = torch.randn(B, R, R, cin, device=dev).to(dt)
    w = (torch.randn(cout, taps * cin, device=dev) * 0.03).to(dt)
    res = torch.randn(B * R * R, cout, device=dev).to(dt)
    mod = torch.rand(B, cout, device=dev) + 0.5
    outs = [torch.empty(B * R * R, cout, dtype=dt, device=dev) for _ in range(3)]
    fullrow = cout == bn and cout <= 256
    for epi in epis:
        d = L.ConvDesc(x=x.data_ptr(), w=w.data_ptr(), B=B, H=R, W=R, cin_pad=cin, cin2_pad=0, cout_pad=cout, taps=taps,
                       block_n=bn, epi_mode=0, flags=0, res_mode=0, res_t=0.3, clip=256.0, tune=int(os.environ.get("VB_TUNE", "0")))
        kinds = [L.VB_OUT_RAW]
        if epi == "qkv":
            if taps != 1 or cout % 192:
                continue
            heads = cout // 192
            qkv = [torch.empty(B * heads * R * R, 64, dtype=dt, device=dev) for _ in range(3)]
            d.epi_mode, d.head_dim, d.parts, d.seg_div = L.VB_EPI_QKVNORM, 64, 3, 1
            for j in range(3):
                d.part_out[j], d.part_seq[j], d.part_off[j] = qkv[j].data_ptr(), R * R, 0
            kinds = []
        elif epi == "mod":
            d.flags, d.mod, d.mod_stride = L.VB_F_MODSILU, mod.data_ptr(), cout
        elif epi != "simple":
            d.flags, d.res = L.VB_F_CLIP, res.data_ptr()
            d.res_mode = L.VB_RES_PIXNORM if (epi.startswith("r2") and fullrow) else L.VB_RES_PLAIN
            if epi.startswith("r3") and fullrow:          # the plans' form: residual scaled by the producer's per-pixel 1/rms side channel
                d.res_mode = L.VB_RES_SCALED
                rn = torch.rand(B * R * R, device=dev) + 0.5
                d.res_rnorm = rn.data_ptr()
                d.flags |= L.VB_F_RESB_FOLDED
            for ch in epi[2:]:
                kinds.append(L.VB_OUT_SILU if ch == "s" else (L.VB_OUT_NORM_SILU if fullrow else L.VB_OUT_SILU))
                if ch == "n":
                    break
            if epi.endswith("nss"):
                kinds.append(L.VB_OUT_SILU)
        for i, k in enumerate(kinds):
            d.out[i], d.out_kind[i], d.out_scale[i] = outs[i].data_ptr(), k, 1.0
        plan = C.c_void_p()
        L.check(lib.vb_plan_create(C.byref(plan)), "create")
        if lib.vb_plan_add_conv(plan, C.byref(d)) != 0:
            print(f"[{idx}] {R}x{R} cin{cin} cout{cout} taps{taps} bn{bn} B{B} {epi:6s}: n/a ({lib.vb_last_error().decode()})", flush=True)
            lib.vb_plan_destroy(plan)
            continue
        for _ in range(3):
            L.check(lib.vb_plan_run(plan, 0, -1, stream), "run")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            L.check(lib.vb_plan_run(plan, 0, -1, stream), "run")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        cyc = C.c_ulonglong(0)
        if int(os.environ.get("VB_DBG", "0")) & 16:
            L.check(lib.vb_debug_conv_cycles(C.byref(cyc)), "cycles")
        stamps = ""
        if int(os.environ.get("VB_DBG", "0")) & 32:
            ts = (C.c_longlong * 8)()
            L.check(lib.vb_debug_conv_stamps(ts), "stamps")
            stamps = "  stamps(cyc): " + " ".join(str(ts[i] - ts[0]) for i in range(1, 7))
        fl = 2.0 * B * R * R * cout * cin * taps
        by = 2.0 * B * R * R * (cin + cout * (len(kinds) + (d.res_mode != 0)))
        print(f"[{idx}] {R}x{R} cin{cin} cout{cout} taps{taps} bn{bn} B{B} {epi:6s}: {ms*1e3:8.1f} us  "
              f"{fl/ms/1e9:7.1f} TFLOP/s  {by/ms/1e6:6.0f} GB/s" +
              (f"  {cyc.value} cyc -> {cyc.value/ms/1e3:.0f} MHz" if cyc.value else "") + stamps, flush=True)
        lib.vb_plan_destroy(plan)
