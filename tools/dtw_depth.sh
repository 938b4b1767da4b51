#!/bin/bash
# pipelining depth of the strong-scaling DTW step (32 pairs per GPU): python bench.py at depths 2, 3, 4
for d in 1 2 3 4; do
  if [ $d = 1 ]; then extra="--pairs 32"; export AFS_BENCH_DTW_DEPTH=2; else extra="--pairs 32"; export AFS_BENCH_DTW_DEPTH=$d; fi
  timeout 250 python bench.py --workloads dtw $extra --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null > /tmp/dd.json
  python - <<PY
import json
d=json.loads(open("/tmp/dd.json").read().strip().splitlines()[-1])
print("depth", $d, "value", round(d["value"],1), "ms_per_step", round(d["ms_per_step"],2), "kernel_ms", d["kernel_ms"])
PY
done
