#!/bin/bash
# ncu --set full of the kernels of one design iteration on 2048x1024 (direct launches, TOPO_NO_GRAPH=1):
#   A: the once-per-design-iteration kernels (multigrid set-up, sensitivity, filter, OC, guards)
#   B: one V-cycle + the three fused fine-grid passes of a PCG iteration
# The reports are reduced to CSV on the box (the .ncu-rep files exceed what gpurun carries back).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TOPO_NO_GRAPH=1
T=/tmp/ncu_r2
mkdir -p $T
python scripts/one_step.py 2 > gpurun_out/r2_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'k_stencil_level1|k_cells_level2|k_cells_coarsen|k_cells_to_stencil|k_coarse_invert|k_oc_multi|k_oc_finish|k_oc_prep|k_oc_step|k_filter|k_simp_sens|k_simp_scale|k_build_diag|k_equilibrium|k_dot|k_mask2' \
    -c 40 -o $T/stage python scripts/one_step.py 1 > gpurun_out/r2_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'k_tail2|k_lvl_|k_restrict|k_prolong_add_fine|k_fine' \
    -s 45 -c 15 -o $T/cycle python scripts/one_step.py 1 > gpurun_out/r2_ncu_b.log 2>&1
for n in stage cycle; do
  ncu -i $T/$n.ncu-rep --page raw --csv > gpurun_out/r2_ncu_${n}_raw.csv 2>/dev/null
done
for k in k_stencil_level1 k_cells_level2 k_oc_multi k_filter_sq k_simp_sens_march; do
  ncu -i $T/stage.ncu-rep --page source --csv --kernel-name regex:$k --launch-count 1 > gpurun_out/r2_ncu_src_$k.csv 2>/dev/null
done
for k in k_tail2 k_lvl_down k_lvl_up k_restrict k_prolong_add_fine; do
  ncu -i $T/cycle.ncu-rep --page source --csv --kernel-name regex:$k --launch-count 1 > gpurun_out/r2_ncu_src_$k.csv 2>/dev/null
done
ls -la gpurun_out/r2_ncu_* ; du -sh gpurun_out
