#!/bin/bash
# call 12: N = 256 tiles on the per-tap kernel for g_0 (I2V_TC_WIDE_N), linear kernel with the (feature, 8 rows) warp mapping
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
( time I2V_TC_WIDE_N=1 timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_conv_tc_gpu.py tests/test_model_gpu.py tests/test_full_size_gpu.py -x -q -m gpu ) > $O/c12_tests.log 2>&1
echo "tests (wide N) rc=$?" > $O/c12_status.txt; tail -8 $O/c12_tests.log
run_bench() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-launches $O/c12_dump_$tag.csv > $O/c12_bench_$tag.json 2>> $O/c12_bench.err
  echo "bench $tag rc=$?" >> $O/c12_status.txt
  python - <<PY
import json
d=json.loads(open("$O/c12_bench_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["roofline"]["families"].items()})
PY
}
run_bench n128 I2V_TC_WIDE_N=0
run_bench n256 I2V_TC_WIDE_N=1
run_bench n128b I2V_TC_WIDE_N=0
run_bench n256b I2V_TC_WIDE_N=1
tail -5 $O/c12_bench.err
cat $O/c12_status.txt
