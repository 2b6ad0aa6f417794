#!/bin/bash
# Round-end GPU pass (one B200): parity suite, smoke, bench lines, ncu launch lists and one full capture per kernel.
# ncu reports stay in /tmp on the box (gpurun_out/ is limited to 64 MiB); only the csv summaries come back.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py 2>gpurun_out/bench_c2.err | tail -1 > gpurun_out/r01_bench_c2.json
python bench.py --workload c3 2>gpurun_out/bench_c3.err | tail -1 > gpurun_out/r01_bench_c3.json
python bench.py --impl reference --steps 10 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r01_bench_c2_reference_arm.json
for w in c2 c3; do
  MOLLYB200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_${w}_bench.csv \
    python bench.py --workload $w --steps 10 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list_$w.log 2>&1
done
MOLLYB200_NO_GRAPH=1 ncu --set full --clock-control none -k regex:build_lists_kernel -s 1 -c 1 -f -o /tmp/build_c2 \
  python bench.py --steps 12 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/ncu_build.log 2>&1
ncu -i /tmp/build_c2.ncu-rep --page raw --csv > gpurun_out/build_c2.raw.csv 2>/dev/null
MOLLYB200_NO_GRAPH=1 ncu --set full --import-source on --clock-control none -k regex:brick_force_kernel -s 5 -c 1 -f -o /tmp/force_c2 \
  python bench.py --steps 12 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/ncu_force.log 2>&1
ncu -i /tmp/force_c2.ncu-rep --page raw --csv > gpurun_out/force_c2.raw.csv 2>/dev/null
python - <<'PY'
import json
for w in ("c2", "c3"):
    d = json.load(open(f"gpurun_out/r01_bench_{w}.json"))
    print(w, d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "force_us", d["stage_us"]["force"],
          "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None, d["roofline"]["frac"], d["fp32"]["frac"])
PY
