#!/bin/bash
# call 4: batched residual loads (+200 registers), dual-output modulate, plane/T-walking modulate split: tests, bench, epilogue stamps
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
( time timeout 300 python -m pytest tests -x -q -m gpu --durations=5 ) > $O/c4_tests.log 2>&1
echo "tests rc=$?" > $O/c4_status.txt; tail -12 $O/c4_tests.log
run_bench() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-launches $O/c4_dump_$tag.csv > $O/c4_bench_$tag.json 2>> $O/c4_bench.err
  echo "bench $tag rc=$?" >> $O/c4_status.txt
  python - <<PY
import json
d=json.loads(open("$O/c4_bench_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["roofline"]["families"].items()})
PY
}
run_bench a A=1
run_bench b A=1
timeout 200 python tools/conv_tc_phases.py epi > $O/c4_phases.txt 2>&1
cat $O/c4_phases.txt
tail -5 $O/c4_bench.err
cat $O/c4_status.txt
