#!/bin/bash
# final single-GPU records of the round: GPU test-suite, default bench line, reference arm
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_final_tests.log
cat gpurun_out/r2_final_tests.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
echo "reference exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'], 'late', d['late']['value'], d['late']['cg_iters'][-1])
print('roofline', {k: d['roofline'][k] for k in ('kernel', 'achieved', 'frac', 'traffic', 'share_of_step')})
print('beso', d['beso_1024x512'])
print('strong', {k: v for k, v in d['strong'].items() if k != 'single_gpu_matvec'})
print('cpu', d['cpu_baseline']['value'], d['same_config_sample']['gpu_512x256']['value'])
r = json.loads(open('gpurun_out/r2_bench_reference.json').read().strip().splitlines()[-1])
print('reference arm', r['value'], r['value_sample_mesh'], r['config']['mesh_that_ran'])
PY
