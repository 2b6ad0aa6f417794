#!/bin/bash
mkdir -p gpurun_out
L=$PWD/molly.jl_b200
run() { echo "== $1"; shift; env "$@" timeout 300 python scripts/sweep.py --workload c2 --configs 3,3,2,8 2>&1 | grep -v mbarrier | cut -c1-330; }
for v in fast nolist nogather neither; do run $v MOLLYB200_LIB=$L/libmb_$v.so; done
