#!/bin/bash
# wstack verification: conv tests first (short timeout), then the full GPU suite, then the bench
TAG=${1:-run}
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -x -q -m gpu > gpurun_out/t_conv.log 2>&1
echo "conv tests rc=$?" > gpurun_out/verify_${TAG}.status
tail -5 gpurun_out/t_conv.log
if grep -q "passed" gpurun_out/t_conv.log && ! grep -q "failed" gpurun_out/t_conv.log; then
  timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1
  echo "all tests rc=$?" >> gpurun_out/verify_${TAG}.status
  tail -3 gpurun_out/t_all.log
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-launches gpurun_out/launch_dump_${TAG}.csv > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
  echo "bench rc=$?" >> gpurun_out/verify_${TAG}.status
  cat gpurun_out/bench_${TAG}.json | cut -c1-600
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1
  echo "smoke rc=$?" >> gpurun_out/verify_${TAG}.status
  tail -2 gpurun_out/smoke_${TAG}.log
fi
cat gpurun_out/verify_${TAG}.status
