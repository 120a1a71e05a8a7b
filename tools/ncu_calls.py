"""One eager denoiser call of each preset between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv ...
(the per-launch list profiles/ keeps: every kernel of a vivid-base, vivid-uncond and vivid-sr call at the bench batch)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda")
plans = []
for i, name in enumerate(("vivid-base", "vivid-uncond", "vivid-sr")):
    net = bench.make_net(name, i, dev)
    p = net.plan(B, dev)
    p.run(graph=False)      # warm-up: weights prepared, buffers touched
    plans.append((name, net, p))
torch.cuda.synchronize()
torch.cuda.profiler.start()
for name, net, p in plans:
    p.run(graph=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches:", {name: p.launches for name, net, p in plans})
