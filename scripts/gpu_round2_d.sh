#!/bin/bash
# validation pass: parity suite + C3/C2 sweeps + bench line with extras
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/d_gputests.log 2>&1; tail -4 gpurun_out/d_gputests.log
timeout 600 python scripts/sweep.py --workload c3 --configs 0,0,0,8 > gpurun_out/d_sweep_c3.log 2>&1; cat gpurun_out/d_sweep_c3.log | cut -c1-400
timeout 900 python bench.py --no-cpu-baseline 2>gpurun_out/d_bench.err | tail -1 > gpurun_out/d_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/d_bench.json'))
print('c2', round(d['value'],1), round(d['ms_per_step']*1e3,1), 'us/step', d['stage_us'], 'e2e', round(d['e2e']['value'],1), d['config']['brick_dims'])
for w,x in d.get('workloads',{}).items():
    print(w, x.get('error') or (round(x['value'],1), round(x['ms_per_step']*1e3,1), x['stage_us'], round(x['e2e']['value'],1), x['config']['brick_dims'], x['rebuilds_in_timed_region'], x['steps']))
PY
