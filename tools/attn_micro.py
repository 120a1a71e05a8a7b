"""Micro-benchmark of vb_attn on the attention shapes of the presets (B = VB_B, default 64).  VB_PRE=1: q pre-scaled by
log2(e)/sqrt(D) as the plans pass it; library knobs: VB_ATTN_POLY (0 | 2 | 3), VB_ATTN_TC, VB_ATTN_DBG."""
import ctypes as C
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vivid_b200 import _lib as L  # noqa: E402

lib = L.lib()
dev = torch.device("cuda")
stream = torch.cuda.current_stream().cuda_stream
B = int(os.environ.get("VB_B", "64"))
SHAPES = [(4, 1024, 2048, 64), (4, 1024, 1024, 64), (6, 256, 512, 64), (6, 256, 256, 64), (8, 64, 128, 64), (8, 1024, 2048, 32)]
dt = L.operand_torch_dtype()
only = os.environ.get("VB_ONLY")
for idx, (h, sq, sk, D) in enumerate(SHAPES):
    if only is not None and str(idx) not in only.split(","):
        continue
    def nrm(t):
        return (t / (1e-4 + t.norm(dim=-1, keepdim=True) / math.sqrt(D))).to(dt)
    q = nrm(torch.randn(B, h, sq, D, device=dev))
    k = nrm(torch.randn(B, h, sk, D, device=dev))
    v = nrm(torch.randn(B, h, sk, D, device=dev))
    y = torch.empty(B, sq, h * D, dtype=dt, device=dev)
    pre = int(os.environ.get("VB_PRE", "1"))
    if pre:
        q = (q.float() * (math.log2(math.e) / math.sqrt(D))).to(dt)
    d = L.AttnDesc(q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), y=y.data_ptr(), B=B, heads=h, sq=sq, sk=sk, head_dim=D, zero_keys=0,
                   q_prescaled=pre)
    for _ in range(3):
        L.check(lib.vb_attn(C.byref(d), stream), "attn")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.check(lib.vb_spin(500, stream), "spin")
    e0.record()
    for _ in range(10):
        L.check(lib.vb_attn(C.byref(d), stream), "attn")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 4.0 * B * h * sq * sk * D
    kf, vf = k.float(), v.float()
    w = (q[:2].float() @ kf[:2].transpose(-1, -2) * (math.log(2.0) if pre else 1.0 / math.sqrt(D))).softmax(-1)
    ref = (w @ vf[:2]).permute(0, 2, 1, 3).reshape(2, sq, h * D)
    err = ((y[:2].float() - ref).norm() / ref.norm()).item()
    print(f"h{h} sq{sq} sk{sk} d{D} B{B}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  rel-L2 {err:.2e}", flush=True)
