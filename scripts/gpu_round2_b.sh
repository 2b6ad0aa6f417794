#!/bin/bash
mkdir -p gpurun_out
L=$PWD/molly.jl_b200
run() { echo "== $1"; shift; env "$@" timeout 300 python scripts/sweep.py --workload c2 --configs 0,0,0,8 3,3,2,8 4,3,3,8 2>&1 | grep -v mbarrier | cut -c1-330; }
for sp in 1 2 4 8; do run split$sp MOLLYB200_LIB=$L/libmb_fast.so MOLLYB200_SPLIT_TAIL=$sp; done
