#!/bin/bash
# tile level kernels (TOPO_LVL_TILED=1, default) against the separate transfer kernels (=0): tests, then A/B through bench.py
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests7.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/r2_tests7.log
export AB_KERNELS="k_restrict k_lvl_prolong k_prolong_add_fine"
for l in 1 2 3; do AB_KERNELS="$AB_KERNELS k_lvl_down_L$l k_lvl_up_L$l"; done
for v in TOPO_LVL_TILED=0 TOPO_LVL_TILED=1 TOPO_LVL_TILED=0 TOPO_LVL_TILED=1; do
  echo "== $v"
  env $v timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-slab-simp 2>/dev/null > gpurun_out/r2_ab7_$v.json
  python - gpurun_out/r2_ab7_$v.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d['value'], 2), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 2), d['detail']['cg_iters'][-3:], d['detail']['compliance_first_last'][1])
for r in d['roofline']['kernels']:
    if r['kernel'].startswith(('k_lvl', 'k_restrict', 'k_prolong')):
        print('   ', r['kernel'], r['avg_us'], r['ms_per_step'], r.get('frac'))
PY
done
