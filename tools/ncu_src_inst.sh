#!/bin/bash
# usage: tools/ncu_src_inst.sh <name> <kernel-regex> <skip>: source-level instruction counts of one launch (top 60 SASS lines
# by executed instructions, aggregated per CUDA source line when -lineinfo maps them)
set -u
name=$1; regex=$2; skip=$3
mkdir -p gpurun_out /tmp/ncu
MFKC_BENCH_NO_CPU=1 ncu --set full --import-source on --clock-control none -k "regex:$regex" -s "$skip" -c 1 -o /tmp/ncu/$name -f python bench.py --steps 1 --warmup 1 > /tmp/ncu/$name.log 2>&1
ncu -i /tmp/ncu/$name.ncu-rep --page source --csv --print-source cuda,sass > /tmp/ncu/srcc_$name.csv 2>/dev/null || ncu -i /tmp/ncu/$name.ncu-rep --page source --csv > /tmp/ncu/srcc_$name.csv 2>/dev/null
python - "$name" <<'PY'
import csv, sys, collections
name = sys.argv[1]
rows = list(csv.reader(open('/tmp/ncu/srcc_%s.csv' % name)))
hi = next(i for i, r in enumerate(rows) if '# Samples' in r or 'Source' in r)
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
print(hdr[:12])
ex = ci.get('# Instructions Executed', ci.get('Instructions Executed'))
src = ci.get('Source')
agg = collections.Counter(); tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    try: v = float(r[ex])
    except Exception: continue
    key = r[src].strip()[:120]
    agg[key] += v; tot += v
with open('gpurun_out/ncu_inst_%s.txt' % name, 'w') as f:
    f.write('total warp instructions %d\n' % tot)
    for k, v in agg.most_common(70): f.write('%5.1f%%  %s\n' % (100 * v / tot, k))
PY
head -40 gpurun_out/ncu_inst_$name.txt
