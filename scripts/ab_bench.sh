#!/bin/bash
# A/B of environment knobs through bench.py (CUDA-event timing, 20 design iterations on 2048x1024): one line per variant
# with design iterations/s, ms per step, end-to-end rate and the in-situ kernel times.
for v in "$@"; do
  echo "== $v"
  env $v timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-slab-simp 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d['detail']['kernel_ms_per_step']
print(round(d['value'], 2), round(d['ms_per_step'], 3), round(d['e2e']['value'], 2), {n: k[n] for n in ('k_tail', 'k_oc_multi') if n in k}, d['detail']['cg_iters'][-3:], d['detail']['compliance_first_last'][1])"
done
