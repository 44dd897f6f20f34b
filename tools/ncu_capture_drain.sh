mkdir -p gpurun_out /tmp/ncu
MFKC_NO_RESIZE=1 MFKC_BENCH_NO_CPU=1 ncu --set full --clock-control none -k "regex:drain_skm" -s 1 -c 1 -o /tmp/ncu/d -f python bench.py --steps 1 --warmup 1 > /tmp/ncu/d.log 2>&1
ncu -i /tmp/ncu/d.ncu-rep --page raw --csv > gpurun_out/ncu_raw_drain_skm_full.csv 2>/dev/null
ncu -i /tmp/ncu/d.ncu-rep --page source --csv > /tmp/ncu/src.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/ncu/src.csv')))
hdr=rows[1]; ci={h:i for i,h in enumerate(hdr)}
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data=[]
for r in rows[2:]:
    if len(r)<len(hdr): continue
    try: v=float(r[ci['# Samples']])
    except: continue
    data.append((v,r))
tot=sum(v for v,_ in data)
agg={s:0 for s in stalls}
for v,r in data:
    for s in stalls:
        try: agg[s]+=float(r[ci[s]])
        except: pass
open('gpurun_out/ncu_stalls_drain_skm_full.txt','w').write('samples %d\n'%tot+'\n'.join('%s %.1f%%'%(k,100*v/tot) for k,v in sorted(agg.items(), key=lambda x:-x[1])[:10])+'\n')
PY
cat gpurun_out/ncu_stalls_drain_skm_full.txt
