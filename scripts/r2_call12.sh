#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 160 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-slab-simp 2>/dev/null > gpurun_out/ab_setup.json
python - <<'PY'
import json
d = json.loads(open('gpurun_out/ab_setup.json').read().strip().splitlines()[-1])
print(round(d['value'], 2), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 2), d['detail']['cg_iters'][-3:], d['detail']['compliance_first_last'][1])
for r in d['roofline']['kernels']:
    if r['kernel'].startswith(('k_cells', 'k_stencil', 'k_interior', 'k_coarse', 'k_build')):
        print('   ', r['kernel'], r['launches_per_step'], r['avg_us'], r['ms_per_step'])
PY
