#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the tcgen05 / bulk-copy paths (B200_PROFILING.md):
python tools/sass_summary.py > profiles/sass_summary.txt     (runs cuobjdump -sass on the in-tree library)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "MUFU.EX2"]
lib = os.path.join(ROOT, "snuffy_b200", "libsnuffy_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
counts, order, cur = {}, [], None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    if cur and re.match(r"\s*/\*[0-9a-f]{4,6}\*/", line):
        counts[cur]["instructions"] += 1
        for k in KEYS:
            if k in line:
                counts[cur][k] += 1
print("# SASS mnemonics per kernel of snuffy_b200/libsnuffy_b200.so (cuobjdump -sass, every cubin is sm_100a)")
print("# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D bulk copy),")
print("# UTMALDG = tensor-map TMA load (the row-plane weight-gradient form gemm_tc_kernel<BN, true>), SYNCS = mbarrier operations")
print("kernel," + ",".join(KEYS) + ",instructions")
for mangled, name in zip(order, names):
    c = counts[mangled]
    if any(c[k] for k in KEYS[:6]):
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("snuffy::", "")
        print(short + "," + ",".join(str(c[k]) for k in KEYS) + f",{c['instructions']}")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("ALL KERNELS (%d)," % len(order) + ",".join(str(tot[k]) for k in KEYS) + f",{tot['instructions']}")
