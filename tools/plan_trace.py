"""Dry run of the library's plan recorder (vb_net_plan_trace — no GPU needed): ops, buffers and device bytes of one denoiser plan.
usage: python tools/plan_trace.py [preset=vivid-base] [batch=128] [--dump]      (presets: vivid-base, vivid-uncond, vivid-sr)"""
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vivid_b200  # noqa: E402
from vivid_b200 import netplan  # noqa: E402

PRESETS = {
    "vivid-base": dict(img_resolution=64, img_channels=3, label_dim=20, model_channels=128, extra_attn=1),
    "vivid-uncond": dict(img_resolution=64, img_channels=3, label_dim=20, model_channels=128, extra_attn=1, uncond=True),
    "vivid-sr": dict(img_resolution=256, img_channels=3, label_dim=20, model_channels=64, super_res=True, noisy_sr=0.25),
}
args = [a for a in sys.argv[1:] if not a.startswith("--")]
name = args[0] if args else "vivid-base"
batches = [int(b) for b in (args[1] if len(args) > 1 else "128").split(",")]
net = vivid_b200.NVPrecond(**PRESETS[name]).eval()
for B in batches:
    trace = netplan.trace_library(net, B)
    if "--dump" in sys.argv:
        print(trace, end="")
    kinds = collections.Counter(l.split()[0] for l in trace.splitlines())
    allocs = [int(l.split()[2]) for l in trace.splitlines() if l.startswith("alloc")]
    io = next(l for l in trace.splitlines() if l.startswith("io "))
    total = int(re.search(r"workspace_bytes=(\d+)", io).group(1))
    ops = sum(kinds[k] for k in ("conv", "attn", "eltwise", "embed", "precond_in", "precond_out"))
    print(f"{name} batch {B}: {ops} ops ({kinds['conv']} conv, {kinds['attn']} attention, {kinds['eltwise']} elementwise, "
          f"{kinds['embed']} embed), {kinds['wprep']} prepared weights, {len(allocs)} buffers, {total / 2**30:.2f} GiB of device memory "
          f"(largest buffer {max(allocs) / 2**20:.0f} MiB)")
