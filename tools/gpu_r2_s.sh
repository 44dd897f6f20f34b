#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s_full.log 2>&1; echo "rc=$?" >> gpurun_out/s_full.log; tail -n 3 gpurun_out/s_full.log
MFKC_BENCH_SAMPLES=8 timeout 600 python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/s_cfg3.json 2> gpurun_out/s_cfg3.err; echo "rc=$?" >> gpurun_out/s_cfg3.err
MFKC_FC_NO_BLOOM=1 MFKC_BENCH_SAMPLES=8 timeout 600 python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/s_cfg3_nobloom.json 2> gpurun_out/s_cfg3_nobloom.err
tail -n 2 gpurun_out/s_cfg3.err
