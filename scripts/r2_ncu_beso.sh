#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TOPO_NO_GRAPH=1
T=/tmp/ncu_r2b
mkdir -p $T
python scripts/beso_one.py > gpurun_out/r2_ncu_beso_plain.log 2>&1 &&
ncu --set full --clock-control none \
    -k regex:'k_energy|k_div_scalar|k_beso_|k_sel_|k_mark_nodes|k_del_node|k_eso_mask|k_stress_nodes|k_vonmises|k_island' \
    -c 60 -o $T/beso python scripts/beso_one.py > gpurun_out/r2_ncu_beso.log 2>&1
ncu -i $T/beso.ncu-rep --page raw --csv > gpurun_out/r2_ncu_beso_raw.csv 2>/dev/null
ls -la gpurun_out/r2_ncu_beso*
