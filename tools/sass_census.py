"""SASS census of libvividb200.so (run here, no GPU): per kernel, how many tcgen05 / TMA / TMEM instructions the binary
holds — the proof that the hot kernels are Blackwell-native (mnemonics: /opt/skills/guides/B200_PROFILING.md).
usage: python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vivid_b200", "libvividb200.so")
MN = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "HMMA", "MUFU", "SYNCS"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        per[cur]["_all"] += 1
        if op in MN:
            per[cur][op] += 1
fam = collections.OrderedDict()
for fn, c in per.items():
    k = re.findall(r"[a-z]+(?:_[a-z0-9]+)*_kernel", fn)
    key = k[-1] if k else fn[:40]
    a = fam.setdefault(key, [0, collections.Counter()])
    a[0] += 1
    a[1].update(c)
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} ({os.path.getsize(LIB)} bytes): instruction census per kernel family")
print(f"# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTMALDG/UTMASTG = TMA tensor load/store, LDTM/STTM = tcgen05.ld/st,")
print(f"# HMMA = legacy mma.sync (attn_kernel only: the 8x8 attention level, 9 us per launch — DESIGN.md §4)")
print(f"{'kernel family':28s} {'variants':>8s} {'instr':>8s} " + " ".join(f"{m:>8s}" for m in MN))
tot = collections.Counter()
for key, (n, c) in fam.items():
    print(f"{key:28s} {n:8d} {c['_all']:8d} " + " ".join(f"{c[m]:8d}" for m in MN))
    tot.update(c)
print(f"{'total':28s} {len(per):8d} {tot['_all']:8d} " + " ".join(f"{tot[m]:8d}" for m in MN))
