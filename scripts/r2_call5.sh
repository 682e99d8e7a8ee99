#!/bin/bash
# Session re-entry check: the whole GPU suite + the default bench line on the restored checkpoint.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r2_tests5.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_tests5.log
timeout 600 python bench.py > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
echo "bench exit $?"
tail -3 gpurun_out/r2_tests5.log
