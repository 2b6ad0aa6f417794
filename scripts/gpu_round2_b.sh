#!/bin/bash
# experiments on the persistent force kernel (C2): variants + one ncu capture
mkdir -p gpurun_out
L=$PWD/molly.jl_b200
run() { echo "== $1"; shift; env "$@" timeout 300 python scripts/sweep.py --workload c2 --configs 3,3,2,8 5,3,3,8 2>&1 | grep -v mbarrier | cut -c1-330; }
run fast MOLLYB200_LIB=$L/libmb_fast.so
run noshift MOLLYB200_LIB=$L/libmb_noshift.so
run static MOLLYB200_LIB=$L/libmb_fast.so MOLLYB200_STATIC_SCHED=1
run nbuf2 MOLLYB200_LIB=$L/libmb_fast.so MOLLYB200_NBUF=2
run nbuf1 MOLLYB200_LIB=$L/libmb_fast.so MOLLYB200_NBUF=1
MOLLYB200_LIB=$L/libmb_fast.so MOLLYB200_NO_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:brick_force_kernel -s 5 -c 1 -f -o /tmp/force_c2 \
  python bench.py --brick 3 3 2 --steps 12 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/b_ncu_force.log 2>&1
ncu -i /tmp/force_c2.ncu-rep --page raw --csv > gpurun_out/b_force_c2.raw.csv 2>/dev/null
ncu -i /tmp/force_c2.ncu-rep --page source --csv > gpurun_out/b_force_c2.source.csv 2>/dev/null
ls -la gpurun_out/b_*
