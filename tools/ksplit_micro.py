"""K-split (vb_conv_desc.tune bit 8) against the single-CTA and CTA-pair layouts on the 3x3 convs of the 8x8 level.

Per batch size and layer: best time of each layout over its legal N tiles (what the plan-time tuner would pick), in
microseconds and useful TFLOP/s.  Launches are timed in batches behind a blocker kernel (vb_spin) with CUDA events.
Env: VB_BATCHES (default 8,32,64,128), VB_REPS.
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
st = torch.cuda.current_stream()
stream = st.cuda_stream
LAYERS = [  # label, R, cin, cout, epilogue
    ("8x8 k4608 n512 conv_res1", 8, 512, 512, "r1"),
    ("8x8 k4608 n512 conv_res0", 8, 512, 512, "mod"),
    ("8x8 k9216 n512 conv_res0", 8, 1024, 512, "mod"),
    ("16x16 k3456 n384 conv_res1", 16, 384, 384, "r1"),
]
LAYOUTS = [("single", 1), ("pair", 2), ("ksplit", 256)]
reps = int(os.environ.get("VB_REPS", "20"))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def time_conv(d):
    if lib.vb_conv(C.byref(d), stream) != 0:
        return None
    best = float("inf")
    for _ in range(3):
        L.check(lib.vb_spin(40, stream), "vb_spin")
        e0.record(st)
        for _ in range(reps):
            lib.vb_conv(C.byref(d), stream)
        e1.record(st)
        st.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


for B in [int(v) for v in os.environ.get("VB_BATCHES", "8,32,64,128").split(",")]:
    for label, R, cin, cout, epi in LAYERS:
        dt = L.operand_torch_dtype()
        x = torch.randn(B, R, R, cin, device=dev).to(dt)
        w = (torch.randn(cout, 9 * cin, device=dev) * 0.03).to(dt)
        res = torch.randn(B * R * R, cout, device=dev).to(dt)
        mod = torch.rand(B, cout, device=dev) + 0.5
        out = torch.empty(B * R * R, cout, dtype=dt, device=dev)
        ws = torch.empty(lib.vb_conv_ksplit_ws_bytes(B, R, R, cout) // 4, device=dev)
        flops = 2.0 * B * R * R * cout * 9 * cin
        cells = []
        for name, tune in LAYOUTS:
            best = None
            for bn in (256, 192, 128, 64):
                if cout % bn:
                    continue
                d = L.ConvDesc(x=x.data_ptr(), w=w.data_ptr(), B=B, H=R, W=R, cin_pad=cin, cout_pad=cout, taps=9, block_n=bn,
                               epi_mode=0, res_t=0.3, clip=256.0, tune=tune, ks_ws=ws.data_ptr() if tune & 256 else None)
                if epi == "mod":
                    d.flags, d.mod, d.mod_stride = L.VB_F_MODSILU, mod.data_ptr(), cout
                else:
                    d.flags, d.res, d.res_mode = L.VB_F_CLIP, res.data_ptr(), L.VB_RES_PLAIN
                d.out[0], d.out_kind[0] = out.data_ptr(), L.VB_OUT_RAW
                t = time_conv(d)
                if t is not None and (best is None or t < best[0]):
                    best = (t, bn)
            cells.append(f"{name} {best[0]:7.1f} us bn{best[1]:<3d} {flops / best[0] * 1e-6:7.1f} TF/s" if best else f"{name} n/a")
        print(f"B={B:<4d} {label:28s} " + " | ".join(cells), flush=True)
