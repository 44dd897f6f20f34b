#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/d_full.log 2>&1; echo "full rc=$?" >> gpurun_out/d_full.log
tail -n 4 gpurun_out/d_full.log
export MFKC_BENCH_NO_CPU=1 MFKC_BENCH_NO_INGEST=1
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?" >> gpurun_out/d_bench.err
export MFKC_BENCH_NO_VERIFY=1
for m in 12 13 15; do MFKC_BIN_M=$m timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/d_bench_m$m.json 2> gpurun_out/d_bench_m$m.err; done
MFKC_BIN_SLACK=1.6 timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/d_bench_slack16.json 2> gpurun_out/d_bench_slack16.err
MFKC_BIN_LOAD=0.55 timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/d_bench_load55.json 2> gpurun_out/d_bench_load55.err
MFKC_BIN_LOAD=0.35 timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/d_bench_load35.json 2> gpurun_out/d_bench_load35.err
tail -n 3 gpurun_out/d_bench.err
