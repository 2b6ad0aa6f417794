#!/bin/bash
# multi-GPU pass on one 8-GPU box: decomposition parity tests, then the strong-scaling bench line for N = 2, 4, 8
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
for n in 2 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-cpu-baseline 2>gpurun_out/bench_n$n.err | tail -1 > gpurun_out/r01_bench_c2_n$n.json
  python -c "import json; d=json.load(open('gpurun_out/r01_bench_c2_n$n.json')); print($n, d['value'], round(d['ms_per_step']*1e3,1), 'us/step', d['stage_us'], 'e2e', d['e2e']['value'] if d.get('e2e') else None, d['config']['parallelism'][:90])"
done
