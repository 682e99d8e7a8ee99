#!/bin/bash
# A/B of environment knobs through bench.py (CUDA-event timing, 20 design iterations on 2048x1024): one line per variant
# with design iterations/s, ms per step, end-to-end rate and the in-situ times of the kernels named in $AB_KERNELS.
for v in "$@"; do
  echo "== $v"
  env $v timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-slab-simp 2>/dev/null | python "$(dirname "$0")/ab_parse.py" ${AB_KERNELS:-}
done
