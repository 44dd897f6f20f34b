#!/bin/bash
# Runs on the GPU box: the launch list of one bench run and one `ncu --set full` capture of each dominant kernel.
set -u
mkdir -p gpurun_out /tmp/ncu
MFKC_BENCH_NO_CPU=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
tools/ncu_kernel.sh extract_skm extract_skm_kernel 30
tools/ncu_kernel.sh drain_skm drain_skm 2
tools/ncu_kernel.sh table_scan table_scan 1
tools/ncu_kernel.sh rs_scatter rs_scatter 12
ls -la gpurun_out | head -40
