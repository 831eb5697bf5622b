#!/bin/bash
# call 7: embedder convs on the tensor-core engine (layer1/2 of the InstanceNorm variant), conv_img with one main accumulator
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
( time timeout 200 python -m pytest tests/test_embedder_gpu.py -x -q -m gpu ) > $O/c7_tests_emb.log 2>&1
echo "embedder tests rc=$?" > $O/c7_status.txt; tail -25 $O/c7_tests_emb.log
( time timeout 300 python -m pytest tests -x -q -m gpu --durations=5 --deselect tests/test_embedder_gpu.py ) > $O/c7_tests.log 2>&1
echo "tests rc=$?" >> $O/c7_status.txt; tail -12 $O/c7_tests.log
run_bench() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-launches $O/c7_dump_$tag.csv > $O/c7_bench_$tag.json 2>> $O/c7_bench.err
  echo "bench $tag rc=$?" >> $O/c7_status.txt
  python - <<PY
import json
d=json.loads(open("$O/c7_bench_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["roofline"]["families"].items()})
PY
}
run_bench a A=1
run_bench b A=1
tail -5 $O/c7_bench.err
grep embedder $O/parity_report.jsonl | tail -8
cat $O/c7_status.txt
