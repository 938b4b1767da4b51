"""profiles/sass_<kernel-file>.txt: per kernel of every object in real-time-audio-sync_b200/build/, the SASS mnemonic histogram
(cuobjdump -sass) and the lines that prove which hardware paths are used (tcgen05: UTCHMMA / UTCBAR / LDTM / STTM /
UTCATOMSWS..., TMA: UBLKCP / UTMALDG, mbarrier: SYNCS, FP64 tensor: DMMA).  python tools/sass_summary.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "real-time-audio-sync_b200", "build")
KEY = re.compile(r"\b(UTC[A-Z0-9]*|LDTM|STTM|UBLKCP|UTMA[A-Z]*|SYNCS|DMMA|HMMA|ELECT|NANOSLEEP|REDUX|SHFL|DFMA|MUFU)\b")
for name in sorted(os.listdir(OBJ)):
    if not name.endswith(".o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, name)], capture_output=True, text=True).stdout
    out = ["# cuobjdump -sass of %s (nvcc -gencode arch=compute_100a,code=sm_100a -O3): mnemonic histogram per kernel" % name, ""]
    kernel, hist, keyops = None, None, None

    def flush():
        if kernel is None or not hist:
            return
        total = sum(hist.values())
        out.append("## %s" % kernel)
        out.append("instructions: %d" % total)
        out.append("hardware-path instructions: " + (", ".join("%s x%d" % kv for kv in sorted(keyops.items())) or "none"))
        out.append("top mnemonics: " + ", ".join("%s %d" % kv for kv in hist.most_common(14)))
        out.append("")

    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            kernel = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()[:160]
            hist, keyops = collections.Counter(), collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and hist is not None:
            op = m.group(1)
            hist[op.split(".")[0]] += 1
            k = KEY.search(op)
            if k:
                keyops[op] += 1
    flush()
    with open(os.path.join(ROOT, "profiles", "sass_%s.txt" % name[:-2]), "w") as fh:
        fh.write("\n".join(out) + "\n")
    print(name, "->", "profiles/sass_%s.txt" % name[:-2])
