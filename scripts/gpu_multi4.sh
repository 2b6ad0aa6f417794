#!/bin/bash
# 4-GPU pass: decomposition parity tests (worlds of 2 and 4), then bench lines at N=1,2,4 under the driver's settings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -8 | tee gpurun_out/m4_tests.log
for n in 1 2 4; do
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n"; fi
  timeout 900 $L bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/m4_bench_n${n}_s20.err | tail -1 > gpurun_out/m4_bench_n${n}_s20.json
  if [ $n != 1 ]; then timeout 600 $L bench.py --gpus $n --no-cpu-baseline --no-extra --no-e2e 2>gpurun_out/m4_bench_n${n}_long.err | tail -1 > gpurun_out/m4_bench_n${n}_long.json; fi
  for f in s20 long; do [ -f gpurun_out/m4_bench_n${n}_$f.json ] && python - <<PY
import json
try:
    d=json.load(open("gpurun_out/m4_bench_n${n}_$f.json"))
    print($n, "$f", round(d['value'],1), round(d['ms_per_step']*1e3,1), 'us/step force', round(d['stage_us']['force'],1), 'e2e', round(d['e2e']['value'],1) if d.get('e2e') else None, 'rebuilds', d['rebuilds_in_timed_region'], 'viol', d['violations'])
    for w,x in (d.get('workloads') or {}).items():
        print('   ', w, x.get('error') or (round(x['value'],1), round(x['ms_per_step']*1e3,1), 'us/step', 'force', round(x['stage_us']['force'],1), 'e2e', round(x['e2e']['value'],1) if x.get('e2e') else None, x['config']['parallelism'][:50]))
except Exception as e:
    print($n, "$f", 'FAILED', e); print(open("gpurun_out/m4_bench_n${n}_$f.err").read()[-1500:])
PY
  done
done
