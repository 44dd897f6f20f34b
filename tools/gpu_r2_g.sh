#!/bin/bash
mkdir -p gpurun_out
export MFKC_BENCH_NO_CPU=1 MFKC_BENCH_NO_INGEST=1 MFKC_BENCH_NO_VERIFY=1
timeout 600 python bench.py --steps 9 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?" >> gpurun_out/g_bench.err
MFKC_BENCH_E2E_LANES=2 timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/g_bench_l2.json 2> gpurun_out/g_bench_l2.err
MFKC_BENCH_E2E_LANES=4 timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/g_bench_l4.json 2> gpurun_out/g_bench_l4.err
tail -n 2 gpurun_out/g_bench.err
