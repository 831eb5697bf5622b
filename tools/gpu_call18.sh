#!/bin/bash
# call 18: two small epilogue-latency items behind switches: stacked-form TMEM loads two chunks per wait (I2V_TC_FLAGS bit 1),
# transpose-reduce in the Linear kernel (I2V_LINEAR_BFLY)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
( time I2V_TC_FLAGS=3 I2V_LINEAR_BFLY=1 timeout 200 python -m pytest tests/test_ops_gpu.py tests/test_conv_tc_gpu.py tests/test_flow_gpu.py tests/test_model_gpu.py tests/test_full_size_gpu.py -x -q -m gpu ) > $O/c18_tests.log 2>&1
echo "tests (switches on) rc=$?" > $O/c18_status.txt; tail -4 $O/c18_tests.log
run_bench() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-launches $O/c18_dump_$tag.csv > $O/c18_bench_$tag.json 2>> $O/c18_bench.err
  echo "bench $tag rc=$?" >> $O/c18_status.txt
  python - <<PY
import json
d=json.loads(open("$O/c18_bench_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["roofline"]["families"].items()})
PY
}
run_bench off A=1
run_bench on I2V_TC_FLAGS=3 I2V_LINEAR_BFLY=1
run_bench offb A=1
run_bench onb I2V_TC_FLAGS=3 I2V_LINEAR_BFLY=1
cat $O/c18_status.txt
