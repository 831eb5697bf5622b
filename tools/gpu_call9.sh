#!/bin/bash
# call 9: SPADE modulate passes plane-ordered (maps stay in L2) vs the T-walking kernel; embedder tensor-core threshold 8
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
( time timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_embedder_gpu.py tests/test_full_size_gpu.py tests/test_full_size_configs_gpu.py -x -q -m gpu ) > $O/c9_tests.log 2>&1
echo "tests rc=$?" > $O/c9_status.txt; tail -6 $O/c9_tests.log
run_bench() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-launches $O/c9_dump_$tag.csv > $O/c9_bench_$tag.json 2>> $O/c9_bench.err
  echo "bench $tag rc=$?" >> $O/c9_status.txt
  python - <<PY
import json
d=json.loads(open("$O/c9_bench_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["roofline"]["families"].items()})
PY
}
run_bench plane A=1
run_bench walk8 I2V_MOD_WALK=8
run_bench planeb A=1
tail -5 $O/c9_bench.err
cat $O/c9_status.txt
