"""Map the warp-stall samples of an ncu report (--import-source on) to source lines through nvdisasm -g of a cubin
compiled from the same source.  usage: ncu_lines.py report.ncu-rep file.cu kernel_substr"""
import csv, re, subprocess, sys, os, collections
rep, cu, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
base = int(data[0][ix['Address']], 16)
cubin = "/tmp/_nl.cubin"
subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC",
                "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr", "-cubin", cu, "-o", cubin], check=True)
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cur = None; line_of = {}; infunc = False
for l in sass.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+),', l)
    if m:
        infunc = ksub in m.group(1); continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*)', l)
    if m and cur: line_of[int(m.group(1), 16)] = cur
reasons = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for r in data:
    off = int(r[ix['Address']], 16) - base
    s = int(r[ix['# Samples']] or 0); tot += s
    a = agg[line_of.get(off, ('?', 0))]
    a[0] += s; a[1] += int(r[ix['Instructions Executed']] or 0)
    for k in reasons:
        if r[ix[k]]: a[2][k[6:]] += int(r[ix[k]])
srcs = {}
def text(f, ln):
    if f not in srcs:
        pth = os.path.join(os.path.dirname(cu), f)
        srcs[f] = open(pth).read().splitlines() if os.path.exists(pth) else []
    t = srcs[f]
    return t[ln - 1].strip()[:80] if 0 < ln <= len(t) else ''
print("total samples", tot, " total warp instr", sum(a[1] for a in agg.values()))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]:
    top = ",".join(f"{n}:{c}" for n, c in a[2].most_common(3))
    print(f"{k[0]}:{k[1]:4d} {a[0]:6d} {a[0]/tot*100:5.1f}% ex {a[1]:9d} [{top}] {text(*k)}")
