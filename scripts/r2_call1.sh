#!/bin/bash
# round 2, GPU call 1: default-path parity, experimental schedules (first time on hardware), timings
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300))"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
step tests_gpu 900 python -m pytest tests -x -q -m gpu
TAUDEM_B200_TEST_EXPERIMENTAL=1 step tests_experimental 400 python -m pytest tests/test_gpu_parity.py -q -k 'experimental or outlets'
step modes_4096 300 python scripts/sweep_modes.py 4096 tiles,levels:8,levels:24,levels:64,levels:24+river:64,hybrid,walk,walk+river:64 2
step modes_16384 600 python scripts/sweep_modes.py 16384 tiles,levels:24,levels:48,levels:auto,levels:24+river:64,levels:auto+river:64,hybrid 2
RIVER_DINF=1 step modes_16384_river_dinf 300 python scripts/sweep_modes.py 16384 levels:24,levels:24+river:64 2
TAUDEM_B200_TIMING=2 step modes_16384_phases 300 python scripts/sweep_modes.py 16384 levels:64,levels:24+river:64,levels:auto+river:64 1
step perf_16384 400 python scripts/gpu_perf.py 16384
TAUDEM_B200_FLATS_BATCH=64 step tests_flats_batch 400 python -m pytest tests/test_gpu_parity.py -q -k "golden or live_reference or depression"
TAUDEM_B200_FLATS_BATCH=64 step perf_16384_flats_batch 400 python scripts/gpu_perf.py 16384
grep -h "DIFFERENT\|identical\|passed\|failed\|Error\|error" gpurun_out/*.log | sort | uniq -c | sort -rn | head -40
