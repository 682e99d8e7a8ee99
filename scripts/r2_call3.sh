#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in TOPO_TAIL_DS=0 TOPO_TAIL_DS=1 TOPO_TAIL_DS=2 "TOPO_B200_LIB=$PWD/solidspy-opt_b200/lib/libtopo_b200_t576.so"; do
  echo "== $v"
  env $v timeout 120 python scripts/tail_stamps.py 2048 1024
done > gpurun_out/r2_tail_stamps3.log 2>&1
bash scripts/ab_bench.sh "TOPO_B200_LIB=$PWD/solidspy-opt_b200/lib/libtopo_b200_t576.so" >> gpurun_out/r2_tail_stamps3.log 2>&1
timeout 300 python -m pytest tests/test_gpu_sized.py -m gpu -q -k "config4" > gpurun_out/r2_tests3.log 2>&1
