#!/bin/bash
# call 10: full suite + smoke + bench with the round's final defaults (persistent halo kernel, embedder on tensor cores from 8 tiles up)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
rm -f $O/parity_report.jsonl
( time timeout 400 python -m pytest tests -x -q -m gpu --durations=5 ) > $O/c10_tests.log 2>&1
echo "tests rc=$?" > $O/c10_status.txt; tail -12 $O/c10_tests.log
( timeout 200 python __graft_entry__.py smoke ) > $O/c10_smoke.log 2>&1
echo "smoke rc=$?" >> $O/c10_status.txt; tail -2 $O/c10_smoke.log
run_bench() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-launches $O/c10_dump_$tag.csv > $O/c10_bench_$tag.json 2>> $O/c10_bench.err
  echo "bench $tag rc=$?" >> $O/c10_status.txt
  python - <<PY
import json
d=json.loads(open("$O/c10_bench_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["roofline"]["families"].items()})
PY
}
run_bench a A=1
run_bench b A=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/c10_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/c10_ncu_launches.log 2>&1
echo "ncu launch list rc=$?" >> $O/c10_status.txt
tail -5 $O/c10_bench.err
cat $O/c10_status.txt
