#!/bin/bash
# call 5: persistent halo kernel (tile loop, next-tile prefetch under the epilogue): conv tests first, full suite, A/B bench, stamps
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
( time timeout 150 python -m pytest tests/test_conv_tc_gpu.py -x -q -m gpu ) > $O/c5_tests_conv.log 2>&1
rc=$?; echo "conv tests (persist=1) rc=$rc" > $O/c5_status.txt; tail -15 $O/c5_tests_conv.log
if [ $rc -ne 0 ]; then
  ( I2V_TC_PERSIST=0 timeout 150 python -m pytest tests/test_conv_tc_gpu.py -x -q -m gpu ) > $O/c5_tests_conv_p0.log 2>&1
  echo "conv tests (persist=0) rc=$?" >> $O/c5_status.txt; tail -15 $O/c5_tests_conv_p0.log
  cat $O/c5_status.txt; exit 1
fi
( time timeout 300 python -m pytest tests -x -q -m gpu --durations=5 ) > $O/c5_tests.log 2>&1
echo "tests rc=$?" >> $O/c5_status.txt; tail -12 $O/c5_tests.log
run_bench() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-launches $O/c5_dump_$tag.csv > $O/c5_bench_$tag.json 2>> $O/c5_bench.err
  echo "bench $tag rc=$?" >> $O/c5_status.txt
  python - <<PY
import json
d=json.loads(open("$O/c5_bench_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["roofline"]["families"].items()})
PY
}
run_bench p1 I2V_TC_PERSIST=1
run_bench p0 I2V_TC_PERSIST=0
run_bench p1b I2V_TC_PERSIST=1
timeout 200 python tools/conv_tc_phases.py epi > $O/c5_phases.txt 2>&1
cat $O/c5_phases.txt
tail -5 $O/c5_bench.err
cat $O/c5_status.txt
