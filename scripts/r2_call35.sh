#!/bin/bash
# round 2, GPU call 35: D8 tall tiles (32 x 64, two rows per lane, 13 workers per SM)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'sweep\|visits\|passed\|failed\|metric' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-700))"; }
step tests_gpu_v16 900 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or live or torture or strip or overlapped"
step stats16_16384 300 python scripts/sweep_stats.py 16384 2
step stats16_65536 600 python scripts/sweep_stats.py 65536 2
