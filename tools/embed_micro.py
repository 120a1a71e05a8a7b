"""vb_embed (Fourier features + embedding linears + every block's modulation vector) timed alone at the presets' sizes.
usage: python tools/embed_micro.py [batches=32,64,128]   (VB_MOD_WIDE=0: the 32-row modulation GEMM kernel for every batch)"""
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for B in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "32,64,128").split(",")]:
    for name, cemb, cnoise, total in (("vivid-base unet", 512, 128, 11904), ("vivid-sr unet", 256, 64, 3712)):
        t = [torch.randn(n, device=dev) for n in (B, B * 20, cnoise, cnoise, cemb * cnoise, cemb * 20, total * cemb, B * cemb, B * total)]
        t[0].abs_().add_(0.1)
        d = L.EmbDesc(sigma=t[0].data_ptr(), geom=t[1].data_ptr(), freqs=t[2].data_ptr(), phases=t[3].data_ptr(), w_noise=t[4].data_ptr(),
                      w_label=t[5].data_ptr(), w_mod=t[6].data_ptr(), emb=t[7].data_ptr(), mod=t[8].data_ptr(), B=B, sigma_n=B,
                      sigma_stride=1, cnoise=cnoise, cemb=cemb, label_dim=20, mod_total=total, geom_rows=B, label_balance=0.5,
                      noise_scale=1.0, geom_scale=1.0)
        best = 1e9
        for _ in range(4):
            L.check(lib.vb_spin(60, st.cuda_stream), "vb_spin")
            e0.record(st)
            for _ in range(20):
                L.check(lib.vb_embed(C.byref(d), st.cuda_stream), "vb_embed")
            e1.record(st)
            st.synchronize()
            best = min(best, e0.elapsed_time(e1) / 20 * 1e3)
        fl = 2.0 * B * total * cemb
        print(f"B={B:<4d} {name:16s} cemb {cemb} mod {total}: {best:7.1f} us per vb_embed ({fl / best * 1e-6:5.1f} TFLOP/s fp32 in the modulation GEMM) "
              f"VB_MOD_WIDE={os.environ.get('VB_MOD_WIDE', '1')}", flush=True)
