#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/p_full.log 2>&1; echo "rc=$?" >> gpurun_out/p_full.log; tail -n 3 gpurun_out/p_full.log
export MFKC_BENCH_NO_CPU=1 MFKC_BENCH_NO_INGEST=1 MFKC_BENCH_NO_VERIFY=1
for i in 1 2 3; do timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/p_bench$i.json 2> gpurun_out/p_bench$i.err; done
