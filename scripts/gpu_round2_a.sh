#!/bin/bash
# first GPU pass of the persistent force kernel: smoke, parity suite, brick-shape sweep, bench
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
tail -3 gpurun_out/a_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/a_gputests.log 2>&1; tail -25 gpurun_out/a_gputests.log
timeout 900 python scripts/sweep.py --workload c2 --configs 0,0,0,8 3,3,2,8 3,3,3,8 4,3,3,8 5,3,3,8 4,4,3,8 2,2,2,8 4,3,2,8 > gpurun_out/a_sweep_c2.log 2>&1; cat gpurun_out/a_sweep_c2.log | cut -c1-400
timeout 600 python scripts/sweep.py --workload c3 --configs 0,0,0,8 1,1,1,8 2,1,1,8 2,2,1,8 2,2,2,8 > gpurun_out/a_sweep_c3.log 2>&1; cat gpurun_out/a_sweep_c3.log | cut -c1-400
timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/a_bench.err | tail -1 > gpurun_out/a_bench_c2.json; cut -c1-1500 gpurun_out/a_bench_c2.json
