#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "features" > gpurun_out/h_tests.log 2>&1; tail -n 2 gpurun_out/h_tests.log
MFKC_BENCH_SAMPLES=8 timeout 600 python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/h_cfg3.json 2> gpurun_out/h_cfg3.err; echo "rc=$?" >> gpurun_out/h_cfg3.err
timeout 600 python bench.py --config 4 --reads 40000000 --steps 3 --warmup 2 > gpurun_out/h_cfg4.json 2> gpurun_out/h_cfg4.err; echo "rc=$?" >> gpurun_out/h_cfg4.err
timeout 600 python bench.py --config 5 --reads 5000000 --steps 3 --warmup 2 > gpurun_out/h_cfg5.json 2> gpurun_out/h_cfg5.err; echo "rc=$?" >> gpurun_out/h_cfg5.err
for f in h_cfg3 h_cfg4 h_cfg5; do echo "== $f"; tail -n 3 gpurun_out/$f.err; head -c 1800 gpurun_out/$f.json; echo; done
