"""Per-denoiser-call latency inside edm_sampler at small batch: the whole loop enqueued from C (vb_sample),
the zero-copy Python loop (plan buffers written by vb_heun, constant inputs uploaded once) and the generic loop (host copies +
clone per call), guided base stage and SR stage, 32 Heun steps.  usage: python tools/sampler_latency.py [batches=2,8]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vivid_b200  # noqa: E402
from vivid_b200.synthetic import synth_batch  # noqa: E402

dev = torch.device("cuda")
net, gnet, sr = (bench.make_net(n, i, dev) for i, n in enumerate(("vivid-base", "vivid-uncond", "vivid-sr")))
print("# ms per denoiser call inside edm_sampler (32 Heun steps = 63 calls; guided: net + gnet per call), wall clock around a "
      "synchronised sampler call, best of 3")
print(f"{'stage':22s} {'B':>3s} {'vb_sample (C loop)':>19s} {'zero-copy loop (Python)':>24s} {'generic loop':>13s} {'graph replay only':>18s}")
for B in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "2,8").split(",")]:
    lo, hi = synth_batch(range(B), 64), synth_batch(range(B), 256)
    src, geom = (lo["src_image"] / 127.5 - 1).to(dev), lo["geometry"].to(dev)
    ssrc, sgeom = (hi["src_image"] / 127.5 - 1).to(dev), hi["geometry"].to(dev)
    noise = torch.randn(B, 3, 64, 64, device=dev)
    snoise = torch.randn(B, 3, 256, 256, device=dev)
    cond = torch.randn(B, 3, 256, 256, device=dev).clamp(-1, 1)
    stages = {
        "base guided (w=1.5)": (lambda: vivid_b200.edm_sampler(net, src, noise, labels=geom, gnet=gnet, num_steps=32, guidance=1.5),
                                [net.plan(B, dev), gnet.plan(B, dev)]),
        "sr": (lambda: vivid_b200.edm_sampler(sr, ssrc, snoise, labels=sgeom, gnet=sr, num_steps=32, conditioning_image=cond),
               [sr.plan(B, dev)]),
    }
    for name, (fn, plans) in stages.items():
        res = []
        for bound, c_loop in (("1", "1"), ("1", "0"), ("0", "0")):
            os.environ["VB_BOUND_SAMPLER"] = bound
            os.environ["VB_C_SAMPLER"] = c_loop
            fn()
            best = 1e9
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            res.append(best / 63 * 1e3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            for p in plans:
                p.run(graph=True)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name:22s} {B:3d} {res[0]:19.3f} {res[1]:24.3f} {res[2]:13.3f} {e0.elapsed_time(e1) / 20:18.3f}")
os.environ.pop("VB_BOUND_SAMPLER", None)
os.environ.pop("VB_C_SAMPLER", None)
