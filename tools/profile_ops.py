"""Per-op CUDA-event profile of one denoiser call of each preset (eager replay); writes a CSV under gpurun_out/."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", f"ops_B{B}.csv")
dev = torch.device("cuda")
rows = []
for i, name in enumerate(("vivid-base", "vivid-uncond", "vivid-sr")):
    net = bench.make_net(name, i, dev)
    p = net.plan(B, dev)
    p.run(graph=False)
    torch.cuda.synchronize()
    prof = p.profile(repeats=3)
    for _ in range(2):
        p.run(graph=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        p.run(graph=True)
    e1.record()
    torch.cuda.synchronize()
    graph_ms = e0.elapsed_time(e1) / 5
    tot = sum(r[4] for r in prof)
    fl = sum(r[2] for r in prof)
    print(f"== {name} B={B}: {tot:.2f} ms/call, {fl/tot/1e9:.1f} TFLOP/s overall, {len(prof)} ops, graph replay {graph_ms:.2f} ms/call, "
          f"mem {torch.cuda.memory_allocated()/2**30:.1f} GiB")
    agg = {}
    for kind, label, f, by, ms in prof:
        a = agg.setdefault((kind, label), [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += ms; a[2] += f; a[3] += by
        rows.append((name, kind, label, f, by, ms))
    for (kind, label), (n, ms, f, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        rate = f"{f/ms/1e9:7.1f} TF/s" if f > 0 else f"{by/ms/1e6:7.1f} GB/s"
        print(f"  {ms:8.3f} ms {100*ms/tot:5.1f}%  x{n:<3d} {kind:8s} {label:34s} {rate}")
    del net, p
    torch.cuda.empty_cache()
with open(out_path, "w") as f:
    f.write("net,kind,label,alg_flops,alg_bytes,ms\n")
    for r in rows:
        f.write(",".join(str(x) for x in r) + "\n")
