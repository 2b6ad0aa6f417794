#!/bin/bash
mkdir -p gpurun_out
MOLLYB200_NO_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:brick_force_kernel -s 5 -c 1 -f -o /tmp/force_c3 \
  python bench.py --workload c3 --steps 12 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/ncu_force_c3.log 2>&1
ncu -i /tmp/force_c3.ncu-rep --page raw --csv > gpurun_out/r02_force_c3.raw.csv 2>/dev/null
ncu -i /tmp/force_c3.ncu-rep --page source --csv > gpurun_out/r02_force_c3.source.csv 2>/dev/null
ls -la gpurun_out/r02_force_c3*
