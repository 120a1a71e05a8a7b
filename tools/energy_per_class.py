"""Energy per useful FLOP (or byte) by layer class: every distinct op of a plan is replayed alone in a
tight loop for `secs` seconds while nvidia-smi samples power and SM clock; J/GFLOP = mean power x time / algorithmic FLOPs.
The whole-step figure at the 1 kW cap is (power / useful TFLOP/s); classes above it are where the energy goes.
usage: python tools/energy_per_class.py [preset=vivid-sr] [batch=64] [secs=0.8] [top=16]"""
import collections
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vivid_b200 import _lib as L  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "vivid-sr"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 0.8
top = int(sys.argv[4]) if len(sys.argv) > 4 else 16
dev = torch.device("cuda")
net = bench.make_net(name, {"vivid-base": 0, "vivid-uncond": 1, "vivid-sr": 2}[name], dev)
p = net.plan(B, dev)
lib = L.lib()
p.run(graph=False)
torch.cuda.synchronize()
prof = p.profile(repeats=2)
classes = collections.OrderedDict()
for i, (kind, label, fl, by, ms) in enumerate(prof):
    c = classes.setdefault((kind, label), dict(ops=[], ms=0.0, fl=fl, by=by))
    c["ops"].append(i)
    c["ms"] += ms
order = sorted(classes.items(), key=lambda kv: -kv[1]["ms"])[:top]
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                        stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append((time.time(), l)) for l in proc.stdout], daemon=True).start()
stream = torch.cuda.current_stream().cuda_stream
print(f"# {name} B={B}: each op class replayed alone for {secs} s (power / SM clock: nvidia-smi, 50 ms samples, first 0.25 s dropped)")
print(f"{'class':52s} {'x/call':>6s} {'us/op':>8s} {'TFLOP/s':>8s} {'GB/s':>7s} {'W':>6s} {'MHz':>6s} {'J/TFLOP':>8s} {'nJ/B':>6s}")
for (kind, label), c in order:
    i = c["ops"][0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(50):
            L.check(lib.vb_plan_run(p.handle, i, i + 1, stream), "run")
        n += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    us = e0.elapsed_time(e1) / n * 1e3
    smp = [l.split(",") for t, l in rows if t0 + 0.25 <= t <= t1 and "," in l]
    clk = sorted(float(s[0]) for s in smp)
    pw = [float(s[1]) for s in smp]
    w = sum(pw) / len(pw) if pw else float("nan")
    tf = c["fl"] / us / 1e6
    gbs = c["by"] / us / 1e3
    print(f"{kind + ' ' + label:52s} {len(c['ops']):6d} {us:8.1f} {tf:8.1f} {gbs:7.0f} {w:6.0f} {clk[len(clk) // 2] if clk else 0:6.0f} "
          f"{(w / tf if tf > 0 else float('nan')):8.2f} {w / gbs if gbs > 0 else float('nan'):6.2f}", flush=True)
    time.sleep(0.3)
proc.terminate()
