#!/bin/bash
# call 6b: same captures as call 6, exported to CSV on the box (the pull-back limit is 64 MiB: only the halo .ncu-rep travels)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 620 --csv --log-file $O/c6_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/c6_ncu_launches.log 2>&1
echo "ncu launch list rc=$?" > $O/c6_status.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo --launch-skip 33 -c 7 -o $O/c6_ncu_halo -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/c6_ncu_halo.log 2>&1
echo "ncu halo rc=$?" >> $O/c6_status.txt
ncu -i $O/c6_ncu_halo.ncu-rep --page raw --csv > $O/c6_ncu_halo_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:modulate8_split_kernel --launch-skip 6 -c 6 -o /tmp/c6_ncu_mod -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/c6_ncu_mod.log 2>&1
echo "ncu modulate rc=$?" >> $O/c6_status.txt
ncu -i /tmp/c6_ncu_mod.ncu-rep --page raw --csv > $O/c6_ncu_mod_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_simt --launch-skip 59 -c 14 -o /tmp/c6_ncu_simt -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/c6_ncu_simt.log 2>&1
echo "ncu simt rc=$?" >> $O/c6_status.txt
ncu -i /tmp/c6_ncu_simt.ncu-rep --page raw --csv > $O/c6_ncu_simt_raw.csv 2>/dev/null
ncu -i /tmp/c6_ncu_simt.ncu-rep --page source --csv --kernel-id :::3 > $O/c6_ncu_simt_source_k3.csv 2>/dev/null
ls -la $O | grep c6_; du -sh $O
cat $O/c6_status.txt
