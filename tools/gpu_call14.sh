#!/bin/bash
# call 14: full suite on the round's final tree
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
rm -f $O/parity_report.jsonl
( time timeout 400 python -m pytest tests -x -q -m gpu --durations=5 ) > $O/c14_tests.log 2>&1
echo "tests rc=$?" > $O/c14_status.txt; tail -12 $O/c14_tests.log
( timeout 200 python __graft_entry__.py smoke ) > $O/c14_smoke.log 2>&1
echo "smoke rc=$?" >> $O/c14_status.txt; tail -1 $O/c14_smoke.log
cat $O/c14_status.txt
