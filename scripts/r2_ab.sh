#!/bin/bash
# A/B of environment knobs through bench.py: one block per variant ("A=1,B=2" = several variables) with the headline, the
# PCG counts, and the in-situ time of every multigrid-level / transfer / fine-grid kernel
# usage: bash scripts/r2_ab.sh TAG VARIANT [VARIANT ...]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
for v in "$@"; do
  echo "== $v"
  env $(echo "$v" | tr ',' ' ') timeout 160 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-slab-simp 2>/dev/null > "gpurun_out/ab_${tag}_$v.json"
  python - "gpurun_out/ab_${tag}_$v.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d['value'], 2), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 2), d['detail']['cg_iters'][-3:], d['detail']['compliance_first_last'][1],
      'late', d['late'] and (round(d['late']['value'], 1), d['late']['cg_iters'][-1]))
for r in d['roofline']['kernels']:
    if r['kernel'].startswith(('k_lvl', 'k_restrict', 'k_prolong', 'k_fine', 'k_tail')):
        print('   ', r['kernel'], r['avg_us'], r['ms_per_step'], r.get('frac'))
PY
done
