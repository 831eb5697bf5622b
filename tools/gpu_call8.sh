#!/bin/bash
# call 8: 4-channel T-walking modulate vs the 8-channel one; embedder tensor-core threshold (32 / 8 / 1 tiles)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
( time timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_embedder_gpu.py tests/test_full_size_gpu.py -x -q -m gpu ) > $O/c8_tests.log 2>&1
echo "tests rc=$?" > $O/c8_status.txt; tail -6 $O/c8_tests.log
run_bench() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-launches $O/c8_dump_$tag.csv > $O/c8_bench_$tag.json 2>> $O/c8_bench.err
  echo "bench $tag rc=$?" >> $O/c8_status.txt
  python - <<PY
import json
d=json.loads(open("$O/c8_bench_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["roofline"]["families"].items()})
PY
}
run_bench walk4 A=1
run_bench walk8 I2V_MOD_WALK=8
run_bench ctas8 I2V_EMB_TC_MIN_CTAS=8
run_bench ctas1 I2V_EMB_TC_MIN_CTAS=1
run_bench walk4b A=1
tail -5 $O/c8_bench.err
cat $O/c8_status.txt
