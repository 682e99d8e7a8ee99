#!/bin/bash
# Round 2, first GPU call: hand-over microbenchmark, the GPU test-suite (with the sized parity tests), the bench line,
# the tail variants.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.log 2>&1
(cd scripts/micro && timeout 60 ./cluster_sync 2000) > gpurun_out/r2_cluster_sync.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=20 > gpurun_out/r2_tests1.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_tests1.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.log 2> gpurun_out/r2_bench1.err
bash scripts/ab_bench.sh TOPO_TAIL_DS=0 TOPO_TAIL_DS=1 TOPO_TAIL_DS=2 TOPO_COARSEST_NODES=160 TOPO_TAIL_MAX=40000 > gpurun_out/r2_ab1.log 2>&1
