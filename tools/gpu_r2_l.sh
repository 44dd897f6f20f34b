#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/l_full.log 2>&1; echo "full rc=$?" >> gpurun_out/l_full.log; tail -n 4 gpurun_out/l_full.log
export MFKC_BENCH_NO_CPU=1 MFKC_BENCH_NO_INGEST=1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "bench rc=$?" >> gpurun_out/l_bench.err
tail -n 2 gpurun_out/l_bench.err
