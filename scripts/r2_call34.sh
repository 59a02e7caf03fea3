#!/bin/bash
# round 2, GPU call 34: group mode v2 (fold over the union of the present directions only)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'sweep\|visits\|passed\|failed\|metric' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-700))"; }
step tests_gpu_v15 900 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or live or torture or strip or overlapped"
step stats15_16384 300 python scripts/sweep_stats.py 16384 2
step stats15_65536 600 python scripts/sweep_stats.py 65536 2
