#!/bin/bash
# round 2, GPU call 4: warp sweep v3 (prop table, scheduler words on separate lines, less polling) + ncu of the stencils and the sweep
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-400))"; }
TAUDEM_B200_TIMING=1 step modes_16384_warp 300 python scripts/sweep_modes.py 16384 tiles,warp 2
step modes_16384_warp_notiming 300 python scripts/sweep_modes.py 16384 warp 3
TAUDEM_B200_TIMING=1 step modes_65536_warp 600 python scripts/sweep_modes.py 65536 warp 1
TAUDEM_B200_SWEEP=warp step tests_gpu_warp 600 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or live_reference or row_strip or geographic"
TAUDEM_B200_SWEEP=warp step ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep_warp|k_d8_stencil|k_dinf_stencil|k_deps_d8|k_deps_dinf|k_fill_init" -s 4 -c 8 -f -o gpurun_out/prof_r02a python scripts/prof_kernels.py 8192
grep -h "DIFFERENT\|identical\|passed\|failed\|Error\|error" gpurun_out/*.log | sort | uniq -c | sort -rn | head -40
