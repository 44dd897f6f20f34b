#!/bin/bash
mkdir -p gpurun_out
export MFKC_BENCH_NO_CPU=1 MFKC_BENCH_NO_INGEST=1 MFKC_BENCH_NO_VERIFY=1 MFKC_BENCH_E2E_SERIAL=1
MFKC_RS_MINB=2 timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r_minb2.json 2> gpurun_out/r_minb2.err
MFKC_RS_MINB=4 timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r_minb4.json 2> gpurun_out/r_minb4.err
unset MFKC_BENCH_NO_VERIFY
MFKC_RS_MINB=4 timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r_full.log 2>&1; echo "rc=$?" >> gpurun_out/r_full.log; tail -n 3 gpurun_out/r_full.log
MFKC_RS_MINB=4 timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r_minb4_verify.json 2> gpurun_out/r_minb4_verify.err
