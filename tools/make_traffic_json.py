"""profiles/traffic.json from ncu captures: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the
kernels bench.py reports a roofline for, stamped with the SHA-1 of the kernel sources so that bench.py can tell a stale
figure from a current one.

  python tools/make_traffic_json.py KEY=REPORT[:per_units] ...
  KEY is one of the names in SPEC below; REPORT an .ncu-rep with one `--set full` result of that kernel;
  per_units divides the bytes (e.g. the number of frames of the captured launch for the per-frame chroma figure).
Existing entries for other keys are kept."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = "real-time-audio-sync_b200/csrc/"
SPEC = {
    "dtw256": ("dtw_wavefront_kernel<double>", "256x20000", [CS + "dtw.cu"]),
    "dtw32": ("dtw_wavefront_kernel<double>", "32x20000", [CS + "dtw.cu"]),
    "dtw128": ("dtw_wavefront_kernel<double>", "128x20000", [CS + "dtw.cu"]),
    "dtw64": ("dtw_wavefront_kernel<double>", "64x20000", [CS + "dtw.cu"]),
    "chroma_tc": ("chroma_tc_spectrum_kernel", "per frame", [CS + "chroma_tc.cu", CS + "tc05.cuh"]),
    "chroma_fp32": ("chroma_fast_kernel<17>", "per frame", [CS + "chroma.cu"]),
    "otw": ("otw_step_kernel", "4096 streams c=500 heavy step", [CS + "otw.cu"]),
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def dram_bytes(report, keep_as):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    with open(keep_as, "w") as fh:          # the raw page is what gets committed (the .ncu-rep stays in gpurun_out/)
        fh.write(out)
    rows = list(csv.reader(io.StringIO(out)))
    head, units, vals = rows[0], rows[1], rows[2]
    total = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        k = head.index(name)
        total += float(vals[k].replace(",", "")) * UNIT[units[k]]
    return total, float(vals[head.index("gpu__time_duration.sum")].replace(",", "")), units[head.index("gpu__time_duration.sum")]


def main():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        entries = json.load(open(path))["entries"]
    except Exception:
        entries = []
    for arg in sys.argv[1:]:
        key, rest = arg.split("=", 1)
        per = 1.0
        if ":" in rest:
            rest, per = rest.rsplit(":", 1)
            per = float(per)
        kernel, config, sources = SPEC[key]
        keep = os.path.join(ROOT, "profiles", "ncu_raw_r2_%s.csv" % key)
        total, dur, dur_unit = dram_bytes(rest, keep)
        e = {"kernel": kernel, "config": config, "dram_bytes": total / per, "captured_launch_dram_bytes": total, "per_units": per,
             "captured_launch_duration": "%s %s (under ncu: cold caches, serialised)" % (dur, dur_unit),
             "source": os.path.relpath(keep, ROOT),
             "src_sha1": {rel: hashlib.sha1(open(os.path.join(ROOT, rel), "rb").read()).hexdigest() for rel in sources}}
        entries = [x for x in entries if not (x["kernel"] == kernel and x["config"] == config)] + [e]
        print(key, "->", e["dram_bytes"], "bytes per unit")
    json.dump({"entries": entries}, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
