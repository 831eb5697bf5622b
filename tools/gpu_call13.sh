#!/bin/bash
# call 13: full suite, default bench line (with the CPU baseline leg), reference arm
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
rm -f $O/parity_report.jsonl
( time timeout 400 python -m pytest tests -x -q -m gpu --durations=5 ) > $O/c13_tests.log 2>&1
echo "tests rc=$?" > $O/c13_status.txt; tail -12 $O/c13_tests.log
( time timeout 400 python bench.py ) > $O/c13_bench_default.json 2> $O/c13_bench_default.err
echo "bench default rc=$?" >> $O/c13_status.txt
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/c13_bench_reference.json 2> $O/c13_bench_reference.err
echo "bench reference rc=$?" >> $O/c13_status.txt
python - <<PY
import json
d=json.loads(open("$O/c13_bench_default.json").read().strip().splitlines()[-1])
print("default", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), d["clocks"], d.get("cpu_baseline"), d["roofline"]["frac"])
r=json.loads(open("$O/c13_bench_reference.json").read().strip().splitlines()[-1])
print("reference", r["value"], r["cpu_baseline"], r["wall_s"])
PY
tail -3 $O/c13_bench_default.err $O/c13_bench_reference.err
cat $O/c13_status.txt
