#!/bin/bash
# N GPUs: slab variants, then bench.py exactly as the driver launches it
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1
bash scripts/r2_call9.sh $N p2p,1,0 | head -3 | cut -c1-700
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench N=$N exit $?"
grep "^{" gpurun_out/r2_bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['clocks'])
print(json.dumps(d['strong']))"
tail -3 gpurun_out/r2_bench_n$N.err | cut -c1-300
