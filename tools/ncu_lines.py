"""Aggregate the warp-stall samples of an ncu report per CUDA source line.
ncu's csv source page is SASS-only; the line of every SASS instruction comes from nvdisasm --print-line-info on the
cubin of the same build (instruction order is identical).
usage: python tools/ncu_lines.py report.ncu-rep object.o kernel_substring [top]"""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os

rep, obj, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
# split per function
lines_of = []
cur = None
infn = False
line = None
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        infn = kname in m.group(1)
        if infn:
            lines_of = []
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines_of.append(line)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
data = rows[2:]
isamp = h.index("# Samples")
iex = h.index("Instructions Executed")
print("sass instructions: ncu %d, nvdisasm %d" % (len(data), len(lines_of)))
agg = {}
for k, r in enumerate(data):
    key = lines_of[k] if k < len(lines_of) else None
    a = agg.setdefault(key, [0, 0])
    a[0] += int(r[isamp])
    a[1] += int(r[iex])
tot = sum(a[0] for a in agg.values())
srcs = {}
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if key:
        for root in ("real-time-audio-sync_b200/csrc", "/usr/local/cuda/include"):
            p = os.path.join(root, key[0])
            if os.path.exists(p):
                if p not in srcs:
                    srcs[p] = open(p).read().splitlines()
                if key[1] - 1 < len(srcs[p]):
                    text = srcs[p][key[1] - 1].strip()[:100]
    print("%5.1f%% %7d inst  %s  %s" % (100.0 * a[0] / max(tot, 1), a[1], key, text))
