#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 6 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-900))"; }
TAUDEM_B200_TIMING=1 step hist_16384 300 python scripts/sweep_modes.py 16384 warp 1
TAUDEM_B200_TIMING=1 step hist_65536 600 python scripts/sweep_modes.py 65536 warp 1
