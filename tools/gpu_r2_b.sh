#!/bin/bash
# round 2, second GPU pass: full suite (new multi-GPU logic on logical shards, CUDA IPC between processes), bench with verify, ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/b_full.log 2>&1; echo "full rc=$?" >> gpurun_out/b_full.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?" >> gpurun_out/b_bench.err
export MFKC_BENCH_NO_CPU=1 MFKC_BENCH_NO_INGEST=1 MFKC_BENCH_NO_VERIFY=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/b_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bin_count -c 1 -o gpurun_out/b_prof_bincount -f python bench.py --steps 1 --warmup 1 > gpurun_out/b_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:extract_skm -s 25 -c 1 -o gpurun_out/b_prof_extract -f python bench.py --steps 1 --warmup 1 > gpurun_out/b_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rs_scatter -s 4 -c 1 -o gpurun_out/b_prof_scatter -f python bench.py --steps 1 --warmup 1 > gpurun_out/b_ncu3.log 2>&1
tail -n 4 gpurun_out/b_full.log; tail -c 600 gpurun_out/b_bench.json; ls -la gpurun_out | tail -n 12
