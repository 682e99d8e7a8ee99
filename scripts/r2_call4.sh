#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in TOPO_TAIL_KIND=2 TOPO_TAIL_KIND=0; do
  echo "== $v"
  env $v timeout 120 python scripts/tail_stamps.py 2048 1024
done > gpurun_out/r2_tail_stamps4.log 2>&1
bash scripts/ab_bench.sh TOPO_TAIL_KIND=2 TOPO_TAIL_KIND=0 >> gpurun_out/r2_tail_stamps4.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests4.log 2>&1
