#!/bin/bash
# list-builder profile (ncu source page) + sweep sanity with the fast experiment build
mkdir -p gpurun_out
L=$PWD/molly.jl_b200
MOLLYB200_LIB=$L/libmb_fast.so timeout 300 python scripts/sweep.py --workload c2 --configs 0,0,0,8 3,3,2,8 2>&1 | grep -v mbarrier | cut -c1-330
MOLLYB200_LIB=$L/libmb_fast.so timeout 300 python scripts/sweep.py --workload c3 --configs 0,0,0,8 2,2,1,8 2>&1 | grep -v mbarrier | cut -c1-330
MOLLYB200_LIB=$L/libmb_fast.so MOLLYB200_NO_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:build_lists_kernel -s 1 -c 1 -f -o /tmp/build_c2 \
  python bench.py --workload c2 --no-extra --steps 6 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/c_ncu_build.log 2>&1
ncu -i /tmp/build_c2.ncu-rep --page raw --csv > gpurun_out/c_build_c2.raw.csv 2>/dev/null
ncu -i /tmp/build_c2.ncu-rep --page source --csv > gpurun_out/c_build_c2.source.csv 2>/dev/null
ls -la gpurun_out/c_*
