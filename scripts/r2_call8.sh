#!/bin/bash
# 2 GPUs: the row-slab check (all transports, one-launch exchanges on/off) against the single-GPU path, then 8192x4096 timed
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
DIST_BIG=${2:-8192x4096} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  tests/dist_simp_check.py > gpurun_out/r2_dist_check_n$N.log 2>&1
echo "exit $?"
grep "^{" gpurun_out/r2_dist_check_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d.items(): print(k, v)
"
tail -5 gpurun_out/r2_dist_check_n$N.log | cut -c1-300
