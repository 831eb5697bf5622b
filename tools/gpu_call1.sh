#!/bin/bash
# round-1 re-entry call: full GPU suite, bench, stage-depth A/B of the halo kernel, ncu captures (modulate, flow, launch list)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/c1_smi.txt 2>&1
( time timeout 900 python -m pytest tests -x -q -m gpu --durations=15 ) > $O/c1_tests.log 2>&1
echo "tests rc=$?" > $O/c1_status.txt; tail -25 $O/c1_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-launches $O/c1_launch_dump.csv > $O/c1_bench.json 2> $O/c1_bench.err
echo "bench rc=$?" >> $O/c1_status.txt; cut -c1-400 $O/c1_bench.json
for ms in 3 4; do
  I2V_TC_MIN_STAGES=$ms timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-launches $O/c1_launch_dump_ms$ms.csv > $O/c1_bench_ms$ms.json 2>> $O/c1_bench.err
  echo "bench ms$ms rc=$?" >> $O/c1_status.txt; cut -c1-330 $O/c1_bench_ms$ms.json
done
for ms in 2 3 4; do
  I2V_TC_MIN_STAGES=$ms timeout 200 python tools/conv_tc_phases.py wide >> $O/c1_phases.txt 2>&1
done
cat $O/c1_phases.txt
# accuracy of the deeper-pipeline configurations (conv tests only)
I2V_TC_MIN_STAGES=4 timeout 300 python -m pytest tests/test_conv_tc_gpu.py -x -q -m gpu > $O/c1_tests_ms4.log 2>&1
echo "conv tests ms4 rc=$?" >> $O/c1_status.txt; tail -3 $O/c1_tests_ms4.log
# ncu: modulate8_split (one decoder pass, BAIR micro-batch 16) and the flow kernel, --set full
timeout 400 ncu --set full --clock-control none --import-source on -k regex:modulate8_split --launch-skip 51 -c 17 -o $O/c1_ncu_modulate -f \
  python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline > $O/c1_ncu_modulate.log 2>&1
echo "ncu modulate rc=$?" >> $O/c1_status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flow_kernel --launch-skip 3 -c 1 -o $O/c1_ncu_flow -f \
  python bench.py --steps 1 --warmup 1 --batch 64 --no-cpu-baseline > $O/c1_ncu_flow.log 2>&1
echo "ncu flow rc=$?" >> $O/c1_status.txt
for r in modulate flow; do
  ncu -i $O/c1_ncu_$r.ncu-rep --page raw --csv > $O/c1_ncu_$r.csv 2>/dev/null
done
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/c1_launches_b64.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/c1_ncu_launches.log 2>&1
echo "ncu launch list rc=$?" >> $O/c1_status.txt
ls -la $O | head -40
cat $O/c1_status.txt
