#!/bin/bash
# round 2, GPU call 31: group mode (2-4 chains, eight lanes per cell) for the D-infinity wavefront
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'sweep\|visits\|passed\|failed\|metric' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-700))"; }
step tests_gpu_v14 900 python -m pytest tests/test_gpu_parity.py -x -q
step stats14_16384 300 python scripts/sweep_stats.py 16384 2
step stats14_65536 600 python scripts/sweep_stats.py 65536 2
TAUDEM_B200_TIMING=1 step stats14_65536_t 600 python scripts/sweep_stats.py 65536 1
