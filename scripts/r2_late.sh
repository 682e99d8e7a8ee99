#!/bin/bash
cd "$(dirname "$0")/.."
for v in TOPO_EXTRAP=1.0 TOPO_EXTRAP=0.5 TOPO_EXTRAP=0.8; do
  env $v timeout 200 python scripts/late_probe.py 100 20 2>&1 | tail -1
done
