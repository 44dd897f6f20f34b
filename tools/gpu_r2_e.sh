#!/bin/bash
mkdir -p gpurun_out
export MFKC_BENCH_NO_CPU=1 MFKC_BENCH_NO_INGEST=1 MFKC_BENCH_NO_VERIFY=1
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?" >> gpurun_out/e_bench.err
MFKC_BIN_PER_CTA=100000 timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/e_bench_persist.json 2> gpurun_out/e_bench_persist.err
MFKC_BENCH_E2E_SERIAL=1 timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/e_bench_serial.json 2> gpurun_out/e_bench_serial.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bin_local or config1 or random_vs or logical_shards_bin" > gpurun_out/e_tests.log 2>&1; tail -n 2 gpurun_out/e_tests.log
