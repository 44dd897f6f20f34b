#!/bin/bash
# Runs on the GPU box: ncu --set full on one launch of each dominant kernel of a full-size bench step and
# leaves small CSV summaries (raw metrics) under gpurun_out/; the .ncu-rep files stay on the box.
set -u
mkdir -p gpurun_out /tmp/ncu
cap() {  # name regex skip
  MFKC_BENCH_NO_CPU=1 ncu --set full --clock-control none -k "regex:$2" -s "$3" -c 1 -o /tmp/ncu/$1 -f python bench.py --steps 1 --warmup 1 > /tmp/ncu/$1.log 2>&1
  ncu -i /tmp/ncu/$1.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$1.csv 2>/dev/null
}
cap drain_skm 'drain_skm' 1
cap extract_skm 'extract_skm_kernel' 30
cap table_scan 'table_scan' 1
cap rs_scatter 'rs_scatter' 12
ls -la gpurun_out
