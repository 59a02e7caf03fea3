#!/bin/bash
# round 2, GPU call 12: D-infinity shares on the fly from the sector table (receiver field in the node words), 16 workers per SM
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'sweep\|visits\|passed\|failed' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-900))"; }
step tests_gpu_v8 900 python -m pytest tests/test_gpu_parity.py -x -q
step stats8_16384 300 python scripts/sweep_stats.py 16384 2
step stats8_65536 600 python scripts/sweep_stats.py 65536 2
TAUDEM_B200_TIMING=1 step stats8_65536_t 600 python scripts/sweep_stats.py 65536 1
