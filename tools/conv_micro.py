"""Micro-benchmark of vb_conv on the layer shapes that dominate the step (B=32 unless given)."""
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
    (16, 384, 384, 9, 128), (8, 512, 512, 9, 32), (8, 512, 512, 9, 128), (16, 384, 1152, 1, 192), (256, 128, 64, 1, 64),
]
B = int(os.environ.get("VB_B", "32"))
reps = int(os.environ.get("VB_REPS", "10"))
only = os.environ.get("VB_ONLY")
for idx, (R, cin, cout, taps, bn) in enumerate(SHAPES):
    if only is not None and str(idx) not in only.split(","):
        continue
    dt = L.operand_torch_dtype()
    x = torch.randn(B, R, R, cin, device=dev).to(dt)
    w = (torch.randn(cout, taps * cin, device=dev) * 0.03).to(dt)
    res = torch.randn(B * R * R, cout, device=dev).to(dt)
    o0 = torch.empty(B * R * R, cout, dtype=dt, device=dev)
    o1 = torch.empty(B * R * R, cout, dtype=dt, device=dev)
    fullrow = cout == bn and cout <= 256
    d = L.ConvDesc(x=x.data_ptr(), w=w.data_ptr(), res=res.data_ptr(), B=B, H=R, W=R, cin_pad=cin, cin2_pad=0,
                   cout_pad=cout, taps=taps, block_n=bn, epi_mode=0, flags=L.VB_F_CLIP,
                   res_mode=L.VB_RES_PIXNORM if fullrow else L.VB_RES_PLAIN, res_t=0.3, clip=256.0)
    d.out[0], d.out_kind[0] = o0.data_ptr(), L.VB_OUT_RAW
    d.out[1], d.out_kind[1] = o1.data_ptr(), (L.VB_OUT_NORM_SILU if fullrow else L.VB_OUT_SILU)
    d.out_scale[1] = 1.0
    plan = C.c_void_p()
    L.check(lib.vb_plan_create(C.byref(plan)), "create")
    L.check(lib.vb_plan_add_conv(plan, C.byref(d)), "add")
    for _ in range(3):
        L.check(lib.vb_plan_run(plan, 0, -1, stream), "run")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        L.check(lib.vb_plan_run(plan, 0, -1, stream), "run")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * B * R * R * cout * cin * taps
    print(f"[{idx}] {R}x{R} cin{cin} cout{cout} taps{taps} bn{bn} B{B}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s", flush=True)
    lib.vb_plan_destroy(plan)
