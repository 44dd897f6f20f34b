#!/bin/bash
mkdir -p gpurun_out
export MFKC_BENCH_NO_CPU=1 MFKC_BENCH_NO_INGEST=1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/m_bench1.json 2> gpurun_out/m_bench1.err
export MFKC_BENCH_NO_VERIFY=1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/m_bench2.json 2> gpurun_out/m_bench2.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/m_bench3.json 2> gpurun_out/m_bench3.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/m_tests.log 2>&1; tail -n 2 gpurun_out/m_tests.log
