#!/bin/bash
# A/B of environment knobs: N design iterations on the headline mesh (scripts/gmg_sweep.py: ms per design iteration,
# total PCG iterations, last compliance -- identical across variants that do not change the arithmetic).
N=${1:-23}
run() { echo "== $*"; env "$@" python scripts/gmg_sweep.py $N 2>&1 | tail -1; }
run TOPO_SYM_MIN=999999999
run TOPO_SYM_MIN=100000
run TOPO_SYM_MIN=100000 TOPO_UNFUSED_MIN=100000
run TOPO_SYM_MIN=100000 TOPO_UNFUSED_MIN=30000
