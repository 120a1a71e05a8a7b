"""BASELINE.json configs[4]: one vivid-base denoiser forward at a large batch, production (fp16 tcgen05) path against the
fp32 validation path: time per call (CUDA events, graph replay / eager), algorithmic TFLOP/s, and the rel-L2 between the two.
usage: python tools/denoiser_micro.py [batch=64] [fp32_batch=8]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vivid_b200.synthetic import synth_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B32 = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda")
net = bench.make_net("vivid-base", 0, dev)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for mode, b, reps in (("fp16 tcgen05", B, 10), ("fp32 validation", B32, 3)):
    fp32 = mode.startswith("fp32")
    p = net.plan(b, dev, fp32=fp32)
    ms = timed(lambda: p.run(graph=True), reps)
    fl = p.alg_flops if fp32 else sum(r[2] for r in p.op_info)
    print(f"vivid-base B={b:3d} {mode:16s}: {ms:8.2f} ms/call  {fl / ms / 1e9:7.1f} TFLOP/s  ({fl / b / 1e9:.1f} GFLOP/image)")

batch = synth_batch(range(B32), 64)
src = (batch["src_image"] / 127.5 - 1).to(dev)
tgt = (batch["tgt_image"] / 127.5 - 1).to(dev)
geom = batch["geometry"].to(dev)
x = tgt + 2.0 * torch.randn(tgt.shape, generator=torch.Generator().manual_seed(3)).to(dev)
sigma = torch.full((B32,), 2.0, device=dev)
a = net(src, x, sigma, geom)
b_ = net(src, x, sigma, geom, force_fp32=True)
print(f"rel-L2 fp16 path vs fp32 path (random-init weights, sigma=2): {((a - b_).norm() / b_.norm()).item():.2e}")
