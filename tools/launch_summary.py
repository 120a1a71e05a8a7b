"""Summarise an ncu per-launch CSV (tools/ncu_calls.py under
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv)
into profiles/: a per-kernel table (launches, time share, DRAM bytes) and the JSON bench.py reads for roofline.traffic.

usage: python tools/launch_summary.py gpurun_out/launches_B64.csv 64 profiles/r01_launches_B64"""
import collections
import csv
import json
import re
import sys

src, batch, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
L = collections.OrderedDict()
for r in rows:
    d = L.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if "time" in r["Metric Name"]:
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    else:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    d[r["Metric Name"]] = v


def short(n):
    n = n.split("(")[0]
    n = re.sub(r"<.*", "", n)
    return n.split("::")[-1]


agg = collections.OrderedDict()
for d in L.values():
    a = agg.setdefault(short(d["name"]), [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d["gpu__time_duration.sum"]
    a[2] += d["dram__bytes_read.sum"]
    a[3] += d["dram__bytes_write.sum"]
tot = sum(a[1] for a in agg.values())
lines = [f"# ncu launch list: one eager denoiser call each of vivid-base, vivid-uncond, vivid-sr at batch {batch}",
         f"# ({len(L)} launches; --clock-control none; per-launch times are cold-cache and serialised: compare SHARES)",
         f"{'kernel':28s} {'launches':>8s} {'ms':>9s} {'share':>7s} {'avg us':>8s} {'dram rd GB':>11s} {'dram wr GB':>11s} {'GB/s':>8s}"]
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{n:28s} {a[0]:8d} {a[1]/1e3:9.3f} {100*a[1]/tot:6.1f}% {a[1]/a[0]:8.1f} {a[2]/1e9:11.3f} {a[3]/1e9:11.3f} "
                 f"{(a[2]+a[3])/a[1]/1e3:8.1f}")
lines.append(f"{'total':28s} {len(L):8d} {tot/1e3:9.3f}")
open(out + ".txt", "w").write("\n".join(lines) + "\n")
conv = [a for n, a in agg.items() if "conv_gemm" in n or n == "ConvKernelParams)"]
assert conv, list(agg)
c = conv[0]
json.dump({"batch": batch, "kernel": "conv_gemm_kernel", "launches": c[0], "dram_bytes_per_launch": (c[2] + c[3]) / c[0],
           "dram_read_bytes": c[2], "dram_write_bytes": c[3], "time_share_under_ncu": c[1] / tot,
           "source": src.split("/")[-1] + " (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, one call per net)"},
          open(out + ".json", "w"), indent=1)
# the raw per-launch list, compact
with open(out + ".csv", "w") as f:
    f.write("id,kernel,grid,block,time_us,dram_read_bytes,dram_write_bytes\n")
    for i, d in L.items():
        f.write(f"{i},{short(d['name'])},{d['grid'].replace(',', ' ')},{d['block'].replace(',', ' ')},{d['gpu__time_duration.sum']:.2f},"
                f"{d['dram__bytes_read.sum']:.0f},{d['dram__bytes_write.sum']:.0f}\n")
print("\n".join(lines))
