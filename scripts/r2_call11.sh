#!/bin/bash
# single GPU: A/B check, ncu launch list of one design iteration (round 2 kernels), ncu --set full of one V-cycle + fine passes
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/r2_ab.sh base2 TOPO_LVL_TILED=1 | head -14
export TOPO_NO_GRAPH=1
T=/tmp/ncu_r2b
mkdir -p $T
python scripts/one_step.py 3 > gpurun_out/r2b_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2b_launches.csv python scripts/one_step.py 3 > gpurun_out/r2b_launches.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'k_tail2|k_lvl_|k_restrict|k_prolong_add_fine|k_fine' \
    -s 45 -c 12 -o $T/cycle python scripts/one_step.py 1 > gpurun_out/r2b_ncu_b.log 2>&1
ncu -i $T/cycle.ncu-rep --page raw --csv > gpurun_out/r2b_ncu_cycle_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2b_ncu_cycle_raw.csv | cut -c1-230
wc -l gpurun_out/r2b_launches.csv
