#!/bin/bash
# ticket order / slack sweep of K2 on the 256-pair batch: GCUPS from bench.py, DRAM bytes from ncu
run() {
  echo "== order=$1 slack=$2"
  AFS_DTW_ORDER=$1 AFS_DTW_SLACK=$2 timeout 300 python bench.py --workloads dtw --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value'],1), 'kernel_ms', round(d['kernel_ms']['accumulate'],2))"
  AFS_DTW_ORDER=$1 AFS_DTW_SLACK=$2 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:dtw_wavefront -s 1 -c 1 python tools/profile_dtw_full.py 256 2>&1 | grep -E "dram__" | awk '{printf "%s %s %s  ", $1, $3, $2} END {print ""}'
}
run p 1
run p 2
run p 4
run g16 0
run g32 0
run g64 0
