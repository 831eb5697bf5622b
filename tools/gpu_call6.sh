#!/bin/bash
# call 6: evidence for the persistent halo kernel (launch list + ncu --set full), ncu of the SPADE modulate passes and the
# embedder's SIMT convs (looking for the limiter), conv_img with one main accumulator
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
( time timeout 300 python -m pytest tests -x -q -m gpu --durations=5 ) > $O/c6_tests.log 2>&1
echo "tests rc=$?" > $O/c6_status.txt; tail -6 $O/c6_tests.log
run_bench() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-launches $O/c6_dump_$tag.csv > $O/c6_bench_$tag.json 2>> $O/c6_bench.err
  echo "bench $tag rc=$?" >> $O/c6_status.txt
  python - <<PY
import json
d=json.loads(open("$O/c6_bench_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["roofline"]["families"].items()})
PY
}
run_bench a A=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 620 --csv --log-file $O/c6_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/c6_ncu_launches.log 2>&1
echo "ncu launch list rc=$?" >> $O/c6_status.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo --launch-skip 33 -c 7 -o $O/c6_ncu_halo -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/c6_ncu_halo.log 2>&1
echo "ncu halo rc=$?" >> $O/c6_status.txt
timeout 300 ncu --set full --clock-control none -k regex:modulate8_split_kernel --launch-skip 6 -c 6 -o $O/c6_ncu_mod -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/c6_ncu_mod.log 2>&1
echo "ncu modulate rc=$?" >> $O/c6_status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_simt --launch-skip 59 -c 14 -o $O/c6_ncu_simt -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/c6_ncu_simt.log 2>&1
echo "ncu simt rc=$?" >> $O/c6_status.txt
ls -la $O | grep c6_
cat $O/c6_status.txt
