#!/bin/bash
# run a subset of the GPU tests on the box: bash tools/gpu_test_one.sh <pytest args...>
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1200 python -m pytest "$@" -x -q -m gpu --durations=8 > gpurun_out/t_one.log 2>&1
echo "rc=$?"; tail -40 gpurun_out/t_one.log
grep full_size gpurun_out/parity_report.jsonl | tail -5
