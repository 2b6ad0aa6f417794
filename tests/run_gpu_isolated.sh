#!/bin/bash
# Run every GPU test in its own process with a hard timeout so that one hung kernel cannot hide the
# results of the others. Usage: tests/run_gpu_isolated.sh [per-test-timeout-seconds] [pytest -k expr]
T=${1:-150}
K=${2:-}
cd "$(dirname "$0")/.."
ids=$(python -m pytest tests -m gpu --collect-only -q -p no:cacheprovider ${K:+-k "$K"} 2>/dev/null | grep "::")
pass=0; fail=0
for id in $ids; do
  out=$(timeout $T python -m pytest "$id" -q -s -p no:cacheprovider 2>&1)
  rc=$?
  if [ $rc -eq 0 ]; then pass=$((pass+1)); echo "PASS $id"; echo "$out" | grep -E "^\[|^    virial|rebuild counts" ;
  else fail=$((fail+1)); echo "FAIL(rc=$rc) $id"; echo "$out" | grep -vE "^\s*$" | tail -25; fi
done
echo "SUMMARY pass=$pass fail=$fail"
[ $fail -eq 0 ]
