#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/i_full.log 2>&1; echo "full rc=$?" >> gpurun_out/i_full.log; tail -n 4 gpurun_out/i_full.log
export MFKC_BENCH_NO_CPU=1 MFKC_BENCH_NO_INGEST=1
MFKC_BENCH_NO_VERIFY=1 timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/i_bench_nv.json 2> gpurun_out/i_bench_nv.err
timeout 900 python bench.py --steps 9 --warmup 3 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "bench rc=$?" >> gpurun_out/i_bench.err
MFKC_BENCH_NO_VERIFY=1 MFKC_BIN_SLACK=1.3 timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/i_bench_s13.json 2> gpurun_out/i_bench_s13.err
tail -n 2 gpurun_out/i_bench.err
