#!/bin/bash
# per-kernel launch lists (ncu, stream mode) for C2 and C3
mkdir -p gpurun_out
for w in c2 c3; do
  MOLLYB200_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_${w}.csv \
    python bench.py --workload $w --steps 30 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list_$w.log 2>&1
  python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/r02_launches_${w}.csv")))
hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        try: agg[d['Kernel Name'].split('(')[0][:60]].append(float(d['Metric Value']))
        except: pass
print("== ${w}")
for k,v in sorted(agg.items(), key=lambda x:-sum(x[1])):
    print(f"{k:62s} n={len(v):4d} mean={sum(v)/len(v)/1000:8.2f} us  total={sum(v)/1e6:8.3f} ms")
PY
done
