#!/bin/bash
# round 2, GPU call 2: the warp-per-tile sweep (first time on hardware): parity, timings, 65536^2
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-400))"; }
step modes_4096_warp 300 python scripts/sweep_modes.py 4096 tiles,warp 2
TAUDEM_B200_SWEEP=warp TAUDEM_B200_TEST_EXPERIMENTAL=1 step tests_gpu_warp 900 python -m pytest tests/test_gpu_parity.py -x -q
step modes_16384_warp 300 python scripts/sweep_modes.py 16384 tiles,warp 3
TAUDEM_B200_TIMING=1 step modes_16384_warp_stats 300 python scripts/sweep_modes.py 16384 tiles,warp 1
step bench_65536_warp 900 python bench.py --steps 3 --warmup 2 --e2e-steps 1 --no-cpu --sweep warp
TAUDEM_B200_TIMING=1 step bench_65536_warp_stats 600 python scripts/sweep_modes.py 65536 warp 1
grep -h "DIFFERENT\|identical\|passed\|failed\|Error\|error" gpurun_out/*.log | sort | uniq -c | sort -rn | head -40
