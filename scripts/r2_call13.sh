#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python bench.py --nx ${1:-8192} --ny ${2:-4096} --steps 4 --warmup 3 --no-cpu-baseline --no-slab-simp --no-late 2>gpurun_out/big.err > gpurun_out/ab_big.json
python - <<'PY'
import json
d = json.loads(open('gpurun_out/ab_big.json').read().strip().splitlines()[-1])
print(round(d['value'], 2), round(d['ms_per_step'], 3), d['detail']['cg_iters'], d['detail']['instrumented_pass_ms_per_step'])
for r in d['roofline']['kernels']:
    print('   ', r['kernel'], r['launches_per_step'], r['avg_us'], r['ms_per_step'], r.get('frac'))
PY
tail -3 gpurun_out/big.err
