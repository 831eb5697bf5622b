#!/bin/bash
# call 2: PDL + residual L2 prefetch A/B, full suite under the new defaults, epilogue stamps, ncu source-level capture of conv_1
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
( time timeout 300 python -m pytest tests -x -q -m gpu --durations=8 ) > $O/c2_tests.log 2>&1
echo "tests rc=$?" > $O/c2_status.txt; tail -15 $O/c2_tests.log
run_bench() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-launches $O/c2_dump_$tag.csv > $O/c2_bench_$tag.json 2>> $O/c2_bench.err
  echo "bench $tag rc=$?" >> $O/c2_status.txt
  python - <<PY
import json
d=json.loads(open("$O/c2_bench_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["roofline"]["families"].items()})
PY
}
run_bench pdl1_f1 I2V_PDL=1 I2V_TC_FLAGS=1
run_bench pdl0_f1 I2V_PDL=0 I2V_TC_FLAGS=1
run_bench pdl1_f0 I2V_PDL=1 I2V_TC_FLAGS=0
run_bench pdl0_f0 I2V_PDL=0 I2V_TC_FLAGS=0
run_bench pdl1_f1_b I2V_PDL=1 I2V_TC_FLAGS=1
for f in 0 1; do
  I2V_TC_FLAGS=$f timeout 200 python tools/conv_tc_phases.py epi >> $O/c2_phases.txt 2>&1
done
cat $O/c2_phases.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo --launch-skip 12 -c 4 -o $O/c2_ncu_halo -f \
  python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline > $O/c2_ncu_halo.log 2>&1
echo "ncu halo rc=$?" >> $O/c2_status.txt
ncu -i $O/c2_ncu_halo.ncu-rep --page raw --csv > $O/c2_ncu_halo_raw.csv 2>/dev/null
ncu -i $O/c2_ncu_halo.ncu-rep --page source --csv > $O/c2_ncu_halo_source.csv 2>/dev/null
ls -la $O | grep c2_
cat $O/c2_status.txt
