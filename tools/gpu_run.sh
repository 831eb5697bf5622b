#!/bin/bash
# One parameterised entry for everything run on the GPU box (replaces the per-call scripts of round 1).
#   gpurun --timeout 1500 -- 'bash tools/gpu_run.sh <cmd> [args] [-- <cmd> [args]] ...'
# Commands (outputs go to gpurun_out/, which gpurun merges back):
#   tests [TAG] [pytest args...]     pytest -m gpu (default: the whole suite), log + parity report
#   bench [TAG] [bench.py args...]   one bench line -> bench_TAG.json (+ launch dump)
#   configs [TAG]                    one bench line per BASELINE config + the reference-gpu leg
#   ab TAG "args A" "args B" ...     ABAB of bench.py argument sets in one box session (same power state)
#   ncu-list [TAG] [bench args]      ncu launch list (gpu__time_duration.sum, --clock-control none) of the bench command
#   ncu-full TAG REGEX SKIP COUNT [bench args]   ncu --set full of COUNT launches matching REGEX, raw csv summary
#   sanitize [TAG]                   tools/sanitize.sh (compute-sanitizer memcheck / racecheck / initcheck / synccheck)
#   smoke                            __graft_entry__.smoke()
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out; mkdir -p $O
summ() {  # print the few numbers of a bench line
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    fam = {k: round(v["ms"], 2) for k, v in d.get("roofline", {}).get("families", {}).items()}
    print(sys.argv[2], "value", round(d["value"]), "e2e", round(d.get("e2e", {}).get("value", 0)), "ms", round(d["ms_per_step"], 2),
          "ab", d.get("ab"), "clk", (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons"), "frac",
          round(d.get("roofline", {}).get("frac", 0), 3), fam, "launches", d.get("gpu_launches"))
except Exception as e:
    print(sys.argv[2], "no bench line:", e)
PY
}
run_one() {
  cmd=$1; shift
  case $cmd in
    tests)
      tag=${1:-all}; [ $# -gt 0 ] && shift
      [ $# -eq 0 ] && set -- tests
      ( time timeout 2400 python -m pytest "$@" -x -q -m gpu --durations=10 ) > $O/tests_$tag.log 2>&1
      echo "tests[$tag] rc=$?" | tee -a $O/status.txt; tail -15 $O/tests_$tag.log ;;
    bench)
      tag=${1:-n1}; [ $# -gt 0 ] && shift
      timeout 900 python bench.py --dump-launches $O/launch_dump_$tag.csv "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err
      echo "bench[$tag] rc=$?" | tee -a $O/status.txt; summ $O/bench_$tag.json $tag ;;
    configs)
      tag=${1:-cfg}
      # CPU baseline leg (10-30 s of host time) only on the headline config; REFGPU=1 adds the reference-gpu legs (minutes)
      for c in bair_b64 bair_b6 bair_b1 landscape_b32_fast landscape_b32 dtdb_fire_seq24_b32 iper128_transfer_b64; do
        extra="--no-cpu-baseline"; [ $c == bair_b64 ] && extra=""
        timeout 900 python bench.py --config $c --steps 5 --warmup 3 $extra > $O/bench_${tag}_$c.json 2> $O/bench_${tag}_$c.err
        echo "bench[$c] rc=$?" | tee -a $O/status.txt; summ $O/bench_${tag}_$c.json $c
      done
      for c in bair_b6 bair_b1; do
        timeout 600 python bench.py --config $c --graph 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${tag}_${c}_graph.json 2> $O/bench_${tag}_${c}_graph.err
        echo "bench[$c graph] rc=$?" | tee -a $O/status.txt; summ $O/bench_${tag}_${c}_graph.json ${c}_graph
      done
      [ "${REFGPU:-0}" == "1" ] && for c in bair_b64 landscape_b32 iper128_transfer_b64; do
        timeout 900 python bench.py --impl reference-gpu --config $c --steps 3 --warmup 2 > $O/bench_${tag}_refgpu_$c.json 2> $O/bench_${tag}_refgpu_$c.err
        echo "reference-gpu[$c] rc=$?" | tee -a $O/status.txt; cut -c1-400 $O/bench_${tag}_refgpu_$c.json
      done ;;
    ab)
      tag=$1; shift
      for rep in 1 2; do i=0; for a in "$@"; do i=$((i+1))
        timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $a > $O/ab_${tag}_${i}_$rep.json 2>> $O/ab_$tag.err
        summ $O/ab_${tag}_${i}_$rep.json "[$a]"
      done; done ;;
    ncu-list)
      tag=${1:-n1}; [ $# -gt 0 ] && shift
      timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_$tag.csv \
        python bench.py --steps 1 --warmup 1 --legs 1 --no-cpu-baseline "$@" > $O/ncu_list_$tag.log 2>&1
      echo "ncu-list[$tag] rc=$?" | tee -a $O/status.txt; python tools/summarise_launches.py $O/launches_$tag.csv > $O/launches_${tag}_summary.csv; head -30 $O/launches_${tag}_summary.csv ;;
    ncu-full)
      tag=$1; re=$2; skip=$3; cnt=$4; shift 4
      timeout 1800 ncu --set full --clock-control none --import-source on -k regex:$re --launch-skip $skip -c $cnt -f -o $O/ncu_$tag \
        python bench.py --steps 1 --warmup 1 --legs 1 --no-cpu-baseline "$@" > $O/ncu_full_$tag.log 2>&1
      echo "ncu-full[$tag] rc=$?" | tee -a $O/status.txt
      ncu -i $O/ncu_$tag.ncu-rep --page raw --csv > $O/ncu_${tag}_raw.csv 2>/dev/null; python tools/summarise_ncu.py $O/ncu_${tag}_raw.csv > $O/ncu_${tag}_summary.txt
      # gpurun merges at most 64 MiB back: the report itself (tens of MB) stays on the box unless KEEP_NCU_REP=1
      [ "${KEEP_NCU_REP:-0}" == "1" ] || rm -f $O/ncu_$tag.ncu-rep
      grep -E "^##|gpu__time_duration.sum|pipe_tensor_cycles_active.avg.pct|dram__bytes_(read|write).sum =|xbar2l1tex_read_bytes.sum.per_second" $O/ncu_${tag}_summary.txt | head -80 ;;
    sanitize)
      bash tools/sanitize.sh ${1:-r02} ;;
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
      echo "smoke rc=$?" | tee -a $O/status.txt; tail -3 $O/smoke.log ;;
    *) echo "unknown command $cmd" ;;
  esac
}
args=()
for a in "$@"; do
  if [ "$a" == "--" ]; then run_one "${args[@]}"; args=(); else args+=("$a"); fi
done
[ ${#args[@]} -gt 0 ] && run_one "${args[@]}"
cat $O/status.txt
