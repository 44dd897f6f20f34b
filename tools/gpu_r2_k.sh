#!/bin/bash
mkdir -p gpurun_out
export MFKC_BENCH_NO_CPU=1 MFKC_BENCH_NO_INGEST=1 MFKC_BENCH_NO_VERIFY=1 MFKC_BENCH_E2E_SERIAL=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bin_count -c 1 -o gpurun_out/k_prof_bincount -f python bench.py --steps 1 --warmup 1 > gpurun_out/k_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:extract_skm -s 25 -c 1 -o gpurun_out/k_prof_extract -f python bench.py --steps 1 --warmup 1 > gpurun_out/k_ncu2.log 2>&1
ls -la gpurun_out/k_prof*
