#!/bin/bash
# bench.py at N = 2 exactly as the driver launches it (strong block = 8192x4096 on two row slabs, last key of the line)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench N=$N exit $?"
tail -c 1500 gpurun_out/r2_bench_n$N.json
tail -5 gpurun_out/r2_bench_n$N.err
