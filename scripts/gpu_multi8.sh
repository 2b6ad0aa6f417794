#!/bin/bash
# 8-GPU sanity: the driver's scaling line at N = 8 (C2 decomposed, C3 replicas, C4 decomposed)
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/m8_bench_n8_s20.err | tail -1 > gpurun_out/m8_bench_n8_s20.json
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/m8_bench_n8_s20.json"))
    print(8, "s20", round(d['value'],1), round(d['ms_per_step']*1e3,1), 'us/step force', round(d['stage_us']['force'],1), 'e2e', round(d['e2e']['value'],1) if d.get('e2e') else None, 'rebuilds', d['rebuilds_in_timed_region'], 'viol', d['violations'], d['config']['parallelism'][:80])
    for w,x in (d.get('workloads') or {}).items():
        print('   ', w, x.get('error') or (round(x['value'],1), round(x['ms_per_step']*1e3,1), 'us/step', 'force', round(x['stage_us']['force'],1), 'e2e', round(x['e2e']['value'],1) if x.get('e2e') else None, x['config']['parallelism'][:50]))
except Exception as e:
    print('FAILED', e); print(open("gpurun_out/m8_bench_n8_s20.err").read()[-2500:])
PY
