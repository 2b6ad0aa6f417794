#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "vv or dynamics or andersen or chunked or triclinic or kinetic" 2>&1 | tail -3
timeout 300 python scripts/sweep.py --workload c2 --configs 0,0,0,8 2>&1 | grep -v mbarrier | cut -c1-330
timeout 300 python scripts/sweep.py --workload c3 --configs 0,0,0,8 2>&1 | grep -v mbarrier | cut -c1-330
timeout 600 python bench.py --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2', round(d['value'],1), round(d['ms_per_step']*1e3,1), d['stage_us'], round(d['e2e']['value'],1))"
