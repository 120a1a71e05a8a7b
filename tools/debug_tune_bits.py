"""Are the tuner's layouts bitwise-equivalent?  Same conv, every (block_n, tune) candidate, outputs compared to the first."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vivid_b200 import _lib as L
lib = L.lib(); dev = torch.device("cuda"); stream = torch.cuda.current_stream().cuda_stream
dt = L.operand_torch_dtype()
for (B, R, cin, cout, taps) in [(5, 16, 64, 64, 9), (5, 16, 128, 128, 9), (5, 8, 128, 128, 9), (4, 64, 128, 128, 9), (5, 16, 64, 192, 1)]:
    torch.manual_seed(1)
    x = torch.randn(B, R, R, cin, device=dev).to(dt)
    w = (torch.randn(cout, taps * cin, device=dev) * 0.05).to(dt)
    ref = None
    for bn in (256, 192, 128, 64):
        if cout % bn: continue
        for tune in (0, 1, 2, 4, 5, 6, 8, 9, 10):
            if taps == 1 and tune >= 4: continue
            out = torch.zeros(B * R * R, cout, dtype=dt, device=dev)
            d = L.ConvDesc(x=x.data_ptr(), w=w.data_ptr(), B=B, H=R, W=R, cin_pad=cin, cin2_pad=0, cout_pad=cout, taps=taps,
                           block_n=bn, epi_mode=0, flags=0, res_mode=0, res_t=0.3, clip=0.0, tune=tune)
            d.out[0], d.out_kind[0], d.out_scale[0] = out.data_ptr(), L.VB_OUT_RAW, 1.0
            rc = lib.vb_conv(C.byref(d), stream)
            if rc != 0: continue
            torch.cuda.synchronize()
            if ref is None: ref = out.clone(); continue
            df = (out.float() - ref.float())
            nd = (df != 0).sum().item()
            if nd: print(f"B{B} R{R} cin{cin} cout{cout} taps{taps} bn{bn} tune{tune}: {nd} elements differ, max abs {df.abs().max().item():.3e}, rel L2 {(df.norm()/ref.float().norm()).item():.3e}")
    print(f"B{B} R{R} cin{cin} cout{cout} taps{taps}: done")
