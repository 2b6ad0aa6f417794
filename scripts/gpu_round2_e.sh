#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_pme.py -x -q -m gpu -s -k "energy_conservation" > gpurun_out/e_tests.log 2>&1; grep -E "energy conservation|C5|passed|failed|Error|error" gpurun_out/e_tests.log | head -30
