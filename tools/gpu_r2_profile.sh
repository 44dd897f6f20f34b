#!/bin/bash
# final profiles of round 2 (one GPU): launch list, DRAM traffic per kernel, full sets of the three top kernels
mkdir -p gpurun_out
export MFKC_BENCH_NO_CPU=1 MFKC_BENCH_NO_INGEST=1 MFKC_BENCH_NO_VERIFY=1 MFKC_BENCH_E2E_SERIAL=1
# warm-up step = 20 mark + 20 extract + 1 place + 1 bin_count + 40 sort + 1 records (+ memsets are not kernels) = ~84 launches; skip them
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 90 --csv --log-file gpurun_out/q_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/q_l.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 90 -c 90 --csv --log-file gpurun_out/q_traffic.csv python bench.py --steps 1 --warmup 1 > gpurun_out/q_t.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bin_count -c 1 -o gpurun_out/q_prof_bincount -f python bench.py --steps 1 --warmup 1 > gpurun_out/q_1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:extract_skm -s 25 -c 1 -o gpurun_out/q_prof_extract -f python bench.py --steps 1 --warmup 1 > gpurun_out/q_2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rs_scatter -s 4 -c 1 -o gpurun_out/q_prof_scatter -f python bench.py --steps 1 --warmup 1 > gpurun_out/q_3.log 2>&1
ls -la gpurun_out/q_*
