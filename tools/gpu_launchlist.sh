#!/bin/bash
# per-kernel durations of one forward (first-step double pass excluded: -s skips it)
TAG=${1:-r01c}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/gpu_quick_time.py ${2:-2:1:1:1} > gpurun_out/${TAG}_ncu.log 2>&1
python - <<PY
import csv,re,collections
lines=[l for l in open('gpurun_out/${TAG}_launches.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
agg=collections.OrderedDict()
for row in rows:
    k=row['Kernel Name']; name=k.split('<')[0].replace('void ','').split('(')[0]
    m=re.search(r'>, (?:\(int\))?(\d+), (?:\(int\))?(\d+)>', k)
    if 'egemm' in name and m: name+=f"<mode{m.group(1)},nh{m.group(2)}>"
    if 'egemm' in name or 'vec_b' in name: name+=" grid=" + row.get('Grid Size','?').strip('()').split(',')[0]
    v=float(row['Metric Value']); a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(v[1] for v in agg.values())
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{k:48s} n={n:4d} total_ms={t/1e6:9.3f} share={100*t/tot:5.1f}% avg_us={t/n/1e3:9.1f}")
PY
