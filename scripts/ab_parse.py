"""stdin: a bench.py line; prints rate, ms per step, e2e rate, the per-step times of the kernels named in argv."""
import json
import sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = {r['kernel']: r['ms_per_step'] for r in d['roofline']['kernels']}
want = [n for n in sys.argv[1:] if n in k] or ['k_tail2', 'k_oc_multi']
print(round(d['value'], 2), round(d['ms_per_step'], 3), round(d['e2e']['value'], 2), {n: k[n] for n in want},
      d['detail']['cg_iters'][-3:], d['detail']['compliance_first_last'][1])
