#!/bin/bash
# call 16: halo kernel with the lo words of a stage on their own barrier (hi*hi MMAs start under the stage's own load)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
( time timeout 150 python -m pytest tests/test_conv_tc_gpu.py -x -q -m gpu ) > $O/c16_tests_conv.log 2>&1
rc=$?; echo "conv tests rc=$rc" > $O/c16_status.txt; tail -5 $O/c16_tests_conv.log
if [ $rc -ne 0 ]; then cat $O/c16_status.txt; exit 1; fi
run_bench() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-launches $O/c16_dump_$tag.csv > $O/c16_bench_$tag.json 2>> $O/c16_bench.err
  echo "bench $tag rc=$?" >> $O/c16_status.txt
  python - <<PY
import json
d=json.loads(open("$O/c16_bench_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["roofline"]["families"].items()})
PY
}
run_bench a A=1
run_bench b A=1
timeout 200 python tools/conv_tc_phases.py epi > $O/c16_phases.txt 2>&1
cat $O/c16_phases.txt
tail -3 $O/c16_bench.err
cat $O/c16_status.txt
