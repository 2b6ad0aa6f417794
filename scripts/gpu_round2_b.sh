#!/bin/bash
mkdir -p gpurun_out
L=$PWD/molly.jl_b200
run() { echo "== $1"; shift; env "$@" timeout 300 python scripts/sweep.py --workload c3 --configs 0,0,0,8 2,1,1,8 3,2,2,8 2>&1 | grep -v mbarrier | cut -c1-330; }
for sp in 1 2 3 4; do run split$sp MOLLYB200_LIB=$L/libmb_fast.so MOLLYB200_SPLIT_SMALL=$sp; done
