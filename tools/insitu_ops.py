"""Per-layer kernel durations recorded in place (CUPTI activity records through torch.profiler): each net's plan is replayed
eagerly (one stream, programmatic dependent launch off, so a record covers exactly one kernel's own run time) right after
a run of graph replays that brings the chip to its sustained clocks.  Launch i of a replay is recorded op i of the plan.

Usage: python tools/insitu_ops.py [B] [out.csv]        (writes gpurun_out/insitu_ops_B<B>.csv by default)
"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vivid_b200 import _lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", f"insitu_ops_B{B}.csv")
REPS = 3
dev = torch.device("cuda")
pk = bench.peaks()
rows = []
for i, name in enumerate(("vivid-base", "vivid-uncond", "vivid-sr")):
    net = bench.make_net(name, i, dev)
    p = net.plan(B, dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(40):                     # ~0.5-1.5 s of graph replays: sustained clocks
        p.run(graph=True)
    e0.record()
    for _ in range(5):
        p.run(graph=True)
    e1.record()
    torch.cuda.synchronize()
    graph_ms = e0.elapsed_time(e1) / 5
    prev = L.lib().vb_set_pdl(0)
    per_call = sum(2 if o[0] == "embed" else 1 for o in p.op_info)          # the embedding op is two kernels (emb, mod)
    ms, good = [0.0] * p.num_ops, 0
    try:
        for _ in range(REPS + 2):
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                p.run(graph=False)
                torch.cuda.synchronize()
            evs = []
            for ev in prof.events():
                dur = getattr(ev, "device_time_total", 0)
                tr = getattr(ev, "time_range", None)
                if tr is not None and "_kernel" in ev.name:
                    evs.append((tr.start, dur / 1e3, ev.name))
            evs.sort()
            if len(evs) != per_call:            # a dropped activity record: this replay cannot be attributed
                print(f"  (replay with {len(evs)} records instead of {per_call} skipped)")
                continue
            j = 0
            for i, o in enumerate(p.op_info):
                for _k in range(2 if o[0] == "embed" else 1):
                    ms[i] += evs[j][1]
                    j += 1
            good += 1
            if good == REPS:
                break
    finally:
        L.lib().vb_set_pdl(prev)
    if good == 0:       # every replay lost an activity record (seen on some boxes for the 150-launch SR plan): no table for this net
        print(f"== {name} B={B}: no complete set of activity records in {REPS + 2} replays; graph replay {graph_ms:.2f} ms/call")
        del net, p
        torch.cuda.empty_cache()
        continue
    ms = [m / good for m in ms]
    tot = sum(ms)
    fl = sum(o[2] for o in p.op_info)
    print(f"== {name} B={B}: kernels {tot:.2f} ms/call in situ ({fl/tot/1e9:.1f} TFLOP/s), graph replay {graph_ms:.2f} ms/call, {p.num_ops} ops")
    agg = {}
    for (kind, label, f, by), m in zip(p.op_info, ms):
        a = agg.setdefault((kind, label), [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += m; a[2] += f; a[3] += by
        rows.append((name, kind, label, f, by, m))
    for (kind, label), (n, m, f, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:34]:
        t_fl, t_by = f / (pk["tflops"] * 1e9), by / (pk["gbs"] * 1e6)        # ms at the measured sustained tensor / HBM peaks
        bound = max(t_fl, t_by)
        rate = f"{f/m/1e9:7.1f} TF/s" if f > 0 else f"{by/m/1e6:7.1f} GB/s"
        print(f"  {m:8.3f} ms {100*m/tot:5.1f}%  x{n:<3d} {kind:8s} {label:34s} {rate}  {by/m/1e6:6.0f} GB/s  roofline {bound:6.3f} ms = {bound/m:4.2f}")
    del net, p
    torch.cuda.empty_cache()
with open(out_path, "w") as f:
    f.write("net,kind,label,alg_flops,alg_bytes,ms_insitu\n")
    for r in rows:
        f.write(",".join(str(x) for x in r) + "\n")
