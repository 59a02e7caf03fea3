#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 3 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300))"; }
TAUDEM_B200_SWEEP=warp step ncu_sweep 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep_warp" -s 0 -c 2 -f -o gpurun_out/prof_r02c python scripts/prof_kernels.py 16384
