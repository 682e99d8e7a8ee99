#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sized.py -m gpu -q -k "beso_1024 or config4" > gpurun_out/r2_tests2.log 2>&1
timeout 120 python scripts/tail_stamps.py 2048 1024 > gpurun_out/r2_tail_stamps.log 2>&1
bash scripts/ab_bench.sh TOPO_NU_TAIL_SMALL=6 TOPO_NU_TAIL_SMALL=8 TOPO_NU_TAIL_BIG=1 TOPO_NU_TAIL_BIG=3 "TOPO_NU_TAIL_BIG=3 TOPO_NU_TAIL_SMALL=8" TOPO_TAIL_MAX=3000 > gpurun_out/r2_ab2.log 2>&1
