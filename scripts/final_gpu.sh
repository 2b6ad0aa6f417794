#!/bin/bash
# Round-end GPU pass (one B200): parity suite, smoke, bench lines (ours + reference arm), ncu launch lists and full captures.
# ncu reports stay in /tmp on the box (gpurun_out/ is limited to 64 MiB); only the csv pages come back.
R=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -s 2>&1 | grep -E "\[|passed|failed|error" | tail -60 > gpurun_out/${R}_gputests.log; tail -3 gpurun_out/${R}_gputests.log
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py 2>gpurun_out/${R}_bench.err | tail -1 > gpurun_out/${R}_bench_c2.json
python bench.py --impl reference --steps 10 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${R}_bench_c2_reference_arm.json
for w in c2 c3; do
  MOLLYB200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_launches_${w}_bench.csv \
    python bench.py --workload $w --steps 40 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list_$w.log 2>&1
done
MOLLYB200_NO_GRAPH=1 ncu --set full --import-source on --clock-control none -k regex:build_lists_kernel -s 1 -c 1 -f -o /tmp/build_c2 \
  python bench.py --workload c2 --steps 6 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/ncu_build.log 2>&1
ncu -i /tmp/build_c2.ncu-rep --page raw --csv > gpurun_out/${R}_build_c2.raw.csv 2>/dev/null
MOLLYB200_NO_GRAPH=1 ncu --set full --import-source on --clock-control none -k regex:brick_force_kernel -s 5 -c 1 -f -o /tmp/force_c2 \
  python bench.py --workload c2 --steps 12 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/ncu_force.log 2>&1
ncu -i /tmp/force_c2.ncu-rep --page raw --csv > gpurun_out/${R}_force_c2.raw.csv 2>/dev/null
ncu -i /tmp/force_c2.ncu-rep --page source --csv > gpurun_out/${R}_force_c2.source.csv 2>/dev/null
MOLLYB200_NO_GRAPH=1 ncu --set full --clock-control none -k regex:vv_kick_drift -s 5 -c 1 -f -o /tmp/k1_c2 \
  python bench.py --workload c2 --steps 12 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/ncu_k1.log 2>&1
ncu -i /tmp/k1_c2.ncu-rep --page raw --csv > gpurun_out/${R}_k1_c2.raw.csv 2>/dev/null
python - <<PY
import json
d = json.load(open("gpurun_out/${R}_bench_c2.json"))
print("c2", d["value"], d["unit"], "us/step", d["ms_per_step"] * 1e3, "e2e", d["e2e"]["value"], "force_us", d["stage_us"]["force"],
      "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None, d["roofline"]["frac"], d["fp32"]["frac"])
for w, x in (d.get("workloads") or {}).items():
    print(w, x.get("error") or (x["value"], x["ms_per_step"] * 1e3, x["e2e"]["value"], x["stage_us"]["force"]))
print(open("gpurun_out/${R}_bench_c2_reference_arm.json").read()[:600])
PY
