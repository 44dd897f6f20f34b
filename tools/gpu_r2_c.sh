#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c_full_$i.log 2>&1; echo "rc=$?" >> gpurun_out/c_full_$i.log; tail -n 3 gpurun_out/c_full_$i.log; done
