#!/bin/bash
# usage: tools/ncu_kernel.sh <name> <kernel-regex> <skip> [env VAR=VAL ...]
# One `ncu --set full` capture of one launch inside a full-size bench step (run on the GPU box).
# Leaves gpurun_out/ncu_raw_<name>.csv (all raw metrics), ncu_stalls_<name>.txt (stall mix) and
# ncu_hot_<name>.txt (the 25 source lines with the most samples).
set -u
name=$1; regex=$2; skip=$3; shift 3
mkdir -p gpurun_out /tmp/ncu
env "$@" MFKC_BENCH_NO_CPU=1 ncu --set full --import-source on --clock-control none -k "regex:$regex" -s "$skip" -c 1 -o /tmp/ncu/$name -f python bench.py --steps 1 --warmup 1 > /tmp/ncu/$name.log 2>&1
ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$name.csv 2>/dev/null
ncu -i /tmp/ncu/$name.ncu-rep --page source --csv > /tmp/ncu/src_$name.csv 2>/dev/null
python - "$name" <<'PY'
import csv, sys
name = sys.argv[1]
rows = list(csv.reader(open('/tmp/ncu/src_%s.csv' % name)))
hi = next(i for i, r in enumerate(rows) if '# Samples' in r or 'Source' in r)
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    try: v = float(r[ci['# Samples']])
    except Exception: continue
    data.append((v, r))
tot = sum(v for v, _ in data) or 1
agg = {s: 0.0 for s in stalls}
for v, r in data:
    for s in stalls:
        try: agg[s] += float(r[ci[s]])
        except Exception: pass
with open('gpurun_out/ncu_stalls_%s.txt' % name, 'w') as f:
    f.write('samples %d\n' % tot)
    for k, v in sorted(agg.items(), key=lambda x: -x[1])[:10]: f.write('%s %.1f%%\n' % (k, 100 * v / tot))
src_col = 'Source' if 'Source' in ci else hdr[1]
ex = ci.get('# Instructions Executed', ci.get('Instructions Executed'))
with open('gpurun_out/ncu_hot_%s.txt' % name, 'w') as f:
    for v, r in sorted(data, key=lambda x: -x[0])[:25]:
        top = sorted(((float(r[ci[s]] or 0), s) for s in stalls), reverse=True)[:2]
        f.write('%5.1f%%  inst=%s  %s | %s\n' % (100 * v / tot, r[ex] if ex is not None else '?', r[ci[src_col]].strip()[:110], ' '.join('%s=%d' % (s[6:], x) for x, s in top)))
PY
cat gpurun_out/ncu_stalls_$name.txt; head -12 gpurun_out/ncu_hot_$name.txt
