#!/bin/bash
# A/B of decoder micro-batch / stream settings in ONE box session (same power state): bash tools/bench_ab.sh TAG
TAG=${1:-ab}
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for cfg in "64 1" "32 2" "32 1" "16 2" "64 1" "32 2"; do
  set -- $cfg
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --micro-batch $1 --streams $2 > gpurun_out/bench_${TAG}_mb$1_s$2.json 2>> gpurun_out/bench_${TAG}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_mb$1_s$2.json").read().strip().splitlines()[-1])
print("mb=$1 streams=$2", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
