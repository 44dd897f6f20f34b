#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/e2e_pipeline_timeline.py 6 > gpurun_out/f_tl_full.txt 2>&1
grep -A3 "link lock" gpurun_out/f_tl_full.txt
nvidia-smi topo -m > gpurun_out/f_topo.txt 2>&1; nproc; free -g | head -2
