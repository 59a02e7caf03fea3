#!/bin/bash
# round 2, GPU call 3: warp sweep v2 (cp.async staging, node/ang in shared memory), ring-pipelined stencils
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-400))"; }
step tests_gpu_core 600 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or live_reference or ties"
TAUDEM_B200_TIMING=1 step modes_4096_warp 300 python scripts/sweep_modes.py 4096 tiles,warp 2
TAUDEM_B200_TIMING=1 step modes_16384_warp 300 python scripts/sweep_modes.py 16384 tiles,warp 2
step modes_16384_warp_notiming 300 python scripts/sweep_modes.py 16384 warp 3
step perf_16384 400 python scripts/gpu_perf.py 16384
TAUDEM_B200_TIMING=1 step modes_65536_warp 600 python scripts/sweep_modes.py 65536 warp 1
grep -h "DIFFERENT\|identical\|passed\|failed\|Error\|error" gpurun_out/*.log | sort | uniq -c | sort -rn | head -40
