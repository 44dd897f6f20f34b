#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export MFKC_BENCH_NO_CPU=1 MFKC_BENCH_NO_INGEST=1
timeout 600 $TR --nproc-per-node 2 --master-port 29531 bench.py --gpus 2 --steps 9 --warmup 3 > gpurun_out/o_n2.json 2> gpurun_out/o_n2.err; echo "rc=$?" >> gpurun_out/o_n2.err
export MFKC_BENCH_NO_VERIFY=1
MFKC_BENCH_E2E_SERIAL=1 timeout 600 $TR --nproc-per-node 2 --master-port 29532 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/o_n2_serial.json 2> gpurun_out/o_n2_serial.err; echo "rc=$?" >> gpurun_out/o_n2_serial.err
MFKC_RS_MINB=3 MFKC_BENCH_E2E_SERIAL=1 timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/o_n1_minb3.json 2> gpurun_out/o_n1_minb3.err
MFKC_RS_MINB=2 MFKC_BENCH_E2E_SERIAL=1 timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/o_n1_minb2.json 2> gpurun_out/o_n1_minb2.err
tail -n 3 gpurun_out/o_n2.err
