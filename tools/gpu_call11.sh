#!/bin/bash
# call 11: linear kernels by shape class (thread-per-feature / warp-per-(feature, 8 rows)), dedicated SPADE 3->128 conv
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
( time timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_flow_gpu.py tests/test_model_gpu.py tests/test_full_size_gpu.py -x -q -m gpu ) > $O/c11_tests.log 2>&1
echo "tests rc=$?" > $O/c11_status.txt; tail -8 $O/c11_tests.log
run_bench() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-launches $O/c11_dump_$tag.csv > $O/c11_bench_$tag.json 2>> $O/c11_bench.err
  echo "bench $tag rc=$?" >> $O/c11_status.txt
  python - <<PY
import json
d=json.loads(open("$O/c11_bench_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["roofline"]["families"].items()})
PY
}
run_bench new A=1
run_bench spadesimt I2V_SPADE_SIMT=1
run_bench newb A=1
tail -5 $O/c11_bench.err
cat $O/c11_status.txt
