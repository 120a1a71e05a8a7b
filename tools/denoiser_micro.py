"""BASELINE.json configs[4]: one vivid-base denoiser forward over a batch sweep, production (fp16 tcgen05) path against the
fp32 validation path: time per call (CUDA events; graph replay with programmatic dependent launch for the fp16 path, eager
for the fp32 one), algorithmic TFLOP/s, fraction of the measured bf16 peak, conv / attention split of the isolated per-op
times, and the rel-L2 between the two paths.
usage: python tools/denoiser_micro.py [fp16 batches, default 1,8,32,64,128] [fp32 batches, default 1,8,32]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vivid_b200.synthetic import synth_batch  # noqa: E402

B16 = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1,8,32,64,128").split(",")]
B32 = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1,8,32").split(",")]
dev = torch.device("cuda")
net = bench.make_net("vivid-base", 0, dev)
pk = bench.peaks()


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


print(f"# vivid-base (250.7 M parameters, 64x64, 161.4 GFLOP/image/call), one denoiser forward; peaks: bf16 sustained "
      f"{pk['tflops']} TFLOP/s, burst {pk['tflops_burst']} ({pk['source']})")
print(f"{'path':16s} {'B':>4s} {'ms/call':>9s} {'TFLOP/s':>9s} {'of burst':>9s} {'images/s':>9s}   conv3 / conv1 / attention TFLOP/s (ops timed alone)")
for b in B16:
    p = net.plan(b, dev)
    reps = max(5, min(200, int(300 / max(0.05 * b, 1))))
    ms = timed(lambda: p.run(graph=True), reps)
    fl = sum(r[2] for r in p.op_info)
    agg = {}
    for kind, label, f, by, op_ms in p.profile(repeats=2):
        a = agg.setdefault(kind, [0.0, 0.0])
        a[0] += f
        a[1] += op_ms
    split = " / ".join(f"{agg[k][0] / agg[k][1] / 1e9:7.1f}" for k in ("conv3", "conv1", "attn"))
    print(f"{'fp16 tcgen05':16s} {b:4d} {ms:9.3f} {fl / ms / 1e9:9.1f} {fl / ms / 1e9 / pk['tflops_burst']:9.3f} {b / ms * 1e3:9.1f}   {split}")
    del p
    net.invalidate_plans()
    torch.cuda.empty_cache()
for b in B32:
    p = net.plan(b, dev, fp32=True)
    ms = timed(lambda: p.run(graph=True), 3 if b <= 8 else 1)
    fl = p.alg_flops
    print(f"{'fp32 validation':16s} {b:4d} {ms:9.2f} {fl / ms / 1e9:9.2f} {fl / ms / 1e9 / pk['tflops_burst']:9.4f} {b / ms * 1e3:9.1f}")
    del p
    net.invalidate_plans()
    torch.cuda.empty_cache()

b = 8
batch = synth_batch(range(b), 64)
src = (batch["src_image"] / 127.5 - 1).to(dev)
tgt = (batch["tgt_image"] / 127.5 - 1).to(dev)
geom = batch["geometry"].to(dev)
x = tgt + 2.0 * torch.randn(tgt.shape, generator=torch.Generator().manual_seed(3)).to(dev)
sigma = torch.full((b,), 2.0, device=dev)
a = net(src, x, sigma, geom)
b_ = net(src, x, sigma, geom, force_fp32=True)
print(f"rel-L2 fp16 path vs fp32 path (random-init weights, sigma=2, B=8): {((a - b_).norm() / b_.norm()).item():.2e}")
