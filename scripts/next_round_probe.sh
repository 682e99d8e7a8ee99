#!/bin/bash
# First GPU call of the next round (nothing here has run yet -- the round-1 GPU budget was spent when it was written):
#   1. scripts/micro/cluster_sync: what one hand-over costs at the tail's launch shape (cluster.sync / relaxed barrier /
#      st.async + mbarrier between neighbours),
#   2. the tail kernel variants on the headline mesh: k_tail (default), k_tail_ds with cluster barriers (=1, measured
#      11 % slower), k_tail_ds with neighbour hand-overs (=2, never run: bounded waits trap instead of hanging),
#   3. the GPU test-suite with the neighbour hand-over, only if step 2 produced a number for it.
# Usage (from the repo root, under gpurun):  bash scripts/next_round_probe.sh > gpurun_out/next_round_probe.log 2>&1
set -u
cd "$(dirname "$0")/.."
( cd scripts/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_sync cluster_sync.cu && timeout 60 ./cluster_sync 2000 )
bash scripts/ab_bench.sh TOPO_TAIL_DS=0 TOPO_TAIL_DS=1 TOPO_TAIL_DS=2
if TOPO_TAIL_DS=2 timeout 60 python scripts/gmg_sweep.py 5 > /dev/null 2>&1; then
  TOPO_TAIL_DS=2 timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
else
  echo "TOPO_TAIL_DS=2 did not complete 5 design iterations: skip the test-suite, look at the trap / the hand-over protocol"
fi
