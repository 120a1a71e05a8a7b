"""Sustained graph replay of one preset (default vivid-sr) for several seconds: ms/call under the power cap, with the
median SM clock and power from nvidia-smi.  usage: python tools/sustained.py [preset] [batch] [seconds]"""
import os, subprocess, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "vivid-sr"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 6.0
dev = torch.device("cuda")
net = bench.make_net(name, 2, dev)
p = net.plan(B, dev)
for _ in range(3):
    p.run(graph=True)
torch.cuda.synchronize()
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                        stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append(l) for l in proc.stdout], daemon=True).start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 0
t0 = time.time()
e0.record()
while time.time() - t0 < secs:
    for _ in range(10):
        p.run(graph=True)
    n += 10
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
proc.terminate()
ms = e0.elapsed_time(e1) / n
clk = sorted(float(r.split(",")[0]) for r in rows[3:] if "," in r)
pw = sorted(float(r.split(",")[1]) for r in rows[3:] if "," in r)
print(f"{name} B={B}: {ms:.2f} ms/call over {n} replays, median SM clock {clk[len(clk)//2] if clk else 0:.0f} MHz, "
      f"median power {pw[len(pw)//2] if pw else 0:.0f} W, env ROWROLL={os.environ.get('VB_ROWROLL','0')} "
      f"AUTOTUNE={os.environ.get('VB_AUTOTUNE','1')} EPI_PP={os.environ.get('VB_EPI_PP','-')} KSPLIT={os.environ.get('VB_KSPLIT','0')} MOD_WIDE={os.environ.get('VB_MOD_WIDE','1')}")
