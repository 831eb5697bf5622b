#!/bin/bash
# compute-sanitizer over the kernel-level GPU tests (small shapes): memcheck, racecheck (shared-memory hazards of the
# mbarrier / TMEM pipelines), initcheck (uninitialised global reads), synccheck.  Logs: gpurun_out/sanitize_<tool>_<TAG>.log
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TAG=${1:-run}; O=gpurun_out; mkdir -p $O
TESTS=${SANITIZE_TESTS:-"tests/test_ops_gpu.py tests/test_flow_gpu.py tests/test_conv_tc_gpu.py"}
# default selection: every kernel family once on small shapes (the full files take > 20 min per tool under the sanitizer)
SEL=${SANITIZE_SELECT:-"not decoder_tensor_core_engine and not full_size and not many_tiles_per_pair and not 130 and not 70 and not 64-20"}
TOOLS=${SANITIZE_TOOLS:-"memcheck racecheck initcheck synccheck"}
for tool in $TOOLS; do
  ( time timeout ${SANITIZE_TIMEOUT:-900} compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest $TESTS -x -q -m gpu -k "$SEL" -p no:cacheprovider ) > $O/sanitize_${tool}_$TAG.log 2>&1
  rc=$?
  echo "sanitize[$tool] rc=$rc $(grep -c 'ERROR SUMMARY' $O/sanitize_${tool}_$TAG.log) summaries: $(grep 'ERROR SUMMARY' $O/sanitize_${tool}_$TAG.log | sort | uniq -c | tr '\n' ';')" | tee -a $O/status.txt
  grep -E "passed|failed|error" $O/sanitize_${tool}_$TAG.log | tail -3
done
