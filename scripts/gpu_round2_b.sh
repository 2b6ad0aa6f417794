#!/bin/bash
mkdir -p gpurun_out
L=$PWD/molly.jl_b200
echo "== packed crf"; MOLLYB200_LIB=$L/libmb_fast.so timeout 300 python scripts/sweep.py --workload c3 --configs 0,0,0,8 2>&1 | grep -v mbarrier | cut -c1-330
echo "== generic"; timeout 300 python scripts/sweep.py --workload c3 --configs 0,0,0,8 2>&1 | grep -v mbarrier | cut -c1-330
MOLLYB200_LIB=$L/libmb_fast.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "6mrr_f32 or all_cut_f32 or dynamics_f32 or molecular_brick" 2>&1 | tail -3
