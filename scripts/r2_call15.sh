#!/bin/bash
# round 2, GPU call 15: wrap sector regular; the full bench line (65536^2 + same-config 16384^2 + CPU sample) and the reference arm
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'sweep\|visits\|passed\|failed\|metric' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-600))"; }
step tests_gpu_v10 900 python -m pytest tests/test_gpu_parity.py -x -q
step stats10_16384 300 python scripts/sweep_stats.py 16384 2
step stats10_65536 600 python scripts/sweep_stats.py 65536 2
step bench_full 1500 python bench.py
step bench_ref 1200 python bench.py --impl reference --steps 1 --warmup 0
