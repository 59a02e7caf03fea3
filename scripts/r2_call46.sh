#!/bin/bash
# round 2, GPU call 46: k_deps_dinf edge path on / off on the 65536^2 bench DEM
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for e in 0 1; do
  TAUDEM_B200_DEPS_EDGE=$e timeout 120 python scripts/deps_dinf_ab.py 65536 5 2>&1 | tail -2 | tee gpurun_out/deps_ab_$e.log
done
