#!/bin/bash
# A/B of kernel build variants: bench.py with MOLLYB200_LIB pointing at each library in turn
for lib in "$@"; do
  for w in c2 c3; do
    MOLLYB200_LIB=$PWD/molly.jl_b200/$lib python bench.py --workload $w --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', '$w', round(d['value'],1), 'steps/s', round(d['ms_per_step']*1e3,1), 'us/step force', round(d['stage_us']['force'],1))"
  done
done
