run() { timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 2000 --warmup 200 --no-cpu-baseline --no-e2e "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1:], d['value'], round(d['ms_per_step']*1e3,1), d['stage_us']['force'], d['rebuilds_in_timed_region'], d['violations'])" "$@"; }
run
run --md-steps-per-call 1000
run --md-steps-per-call 1000 --rebuild-every 200
run --md-steps-per-call 1000 --no-cm
run --md-steps-per-call 1000 --no-cm --rebuild-every 200
