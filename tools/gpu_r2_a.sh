#!/bin/bash
# round 2, first GPU pass: new bin-local path
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bin_local or config1 or tinytest or empty_inputs" > gpurun_out/a_focus.log 2>&1; echo "focus rc=$?" >> gpurun_out/a_focus.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/a_full.log 2>&1; echo "full rc=$?" >> gpurun_out/a_full.log
MFKC_BENCH_NO_CPU=1 MFKC_BENCH_NO_INGEST=1 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err
MFKC_BIN_TMA=0 MFKC_BENCH_NO_CPU=1 MFKC_BENCH_NO_INGEST=1 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench_notma.json 2> gpurun_out/a_bench_notma.err
MFKC_BENCH_VARIANT=table MFKC_BENCH_NO_CPU=1 MFKC_BENCH_NO_INGEST=1 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench_table.json 2> gpurun_out/a_bench_table.err
tail -3 gpurun_out/a_focus.log gpurun_out/a_full.log; tail -c 1500 gpurun_out/a_bench.json
