#!/bin/bash
# A/B of environment knobs on the headline mesh: each line = one process of scripts/gmg_sweep.py (N design
# iterations from rho = 0.5; prints ms per design iteration, total PCG iterations and the last compliance, which
# must be identical across variants that do not change the arithmetic).
N=${1:-23}
run() { echo "== $*"; env "$@" python scripts/gmg_sweep.py $N 2>&1 | tail -1; }
run TOPO_PDL=0
run TOPO_PDL=1
run TOPO_PDL=3
run TOPO_PDL=5
run TOPO_PDL=7
run TOPO_PDL=0
