#!/bin/bash
# call 3: fused shortcut (side input) + T-walking modulate: tests, bench A/B
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
( time timeout 300 python -m pytest tests -x -q -m gpu --durations=5 ) > $O/c3_tests.log 2>&1
echo "tests rc=$?" > $O/c3_status.txt; tail -25 $O/c3_tests.log
run_bench() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-launches $O/c3_dump_$tag.csv > $O/c3_bench_$tag.json 2>> $O/c3_bench.err
  echo "bench $tag rc=$?" >> $O/c3_status.txt
  python - <<PY
import json
d=json.loads(open("$O/c3_bench_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["roofline"]["families"].items()})
PY
}
run_bench fused A=1
run_bench nofuse I2V_NO_FUSE_S=1
run_bench fused_b A=1
run_bench fused_pdl0 I2V_PDL=0
tail -5 $O/c3_bench.err
cat $O/c3_status.txt
