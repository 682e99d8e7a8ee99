#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1; shift
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 \
  scripts/slab_variants.py 8192 4096 "$*" > gpurun_out/slab_variants_n$N.log 2>&1
echo "exit $?"
grep "^{" gpurun_out/slab_variants_n$N.log
tail -3 gpurun_out/slab_variants_n$N.log | cut -c1-300
