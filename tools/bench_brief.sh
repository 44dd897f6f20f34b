#!/bin/bash
# usage: tools/bench_brief.sh [bench args]; prints a one-line summary of bench.py's JSON
MFKC_BENCH_NO_CPU=1 python bench.py "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']
print('value %.2f G/s  step %.1f ms  e2e %.2f G/s | '%(d['value']/1e9,d['ms_per_step'],d['e2e']['value']/1e9)+' '.join('%s=%.1f'%(a,b) for a,b in k.items())+' | frac %.3f'%d['roofline']['frac']+' drains/step %s'%d['roofline']['kernel_launches_per_step'].get('bin_count'))"
