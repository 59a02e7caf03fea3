#!/bin/bash
# round 2, GPU call 5: warp sweep v4 (row ready masks, fork stack, compact D-infinity gather), compact D-infinity stencil, new bench.py
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-400))"; }
TAUDEM_B200_TIMING=1 step modes_16384_warp 300 python scripts/sweep_modes.py 16384 tiles,warp 2
step perf_16384 400 python scripts/gpu_perf.py 16384
TAUDEM_B200_TIMING=1 step modes_65536_warp 600 python scripts/sweep_modes.py 65536 warp 1
TAUDEM_B200_SWEEP=warp step tests_gpu_warp 600 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or live_reference or row_strip or ties"
step bench_ref_4096 600 python bench.py --impl reference --size 4096 --steps 1 --warmup 0 --cpu-ranks 16
step bench_4096 600 python bench.py --size 4096 --steps 3 --warmup 3 --sweep warp --cpu-ranks 16 --cpu-sample 2048
TAUDEM_B200_SWEEP=warp step ncu_full 600 ncu --set full --clock-control none --import-source on -k regex:"k_d8_stencil|k_dinf_stencil|k_fill_init|k_deps_d8" -s 3 -c 4 -f -o gpurun_out/prof_r02b python scripts/prof_kernels.py 8192
grep -h "DIFFERENT\|identical\|passed\|failed\|Error\|error" gpurun_out/*.log | sort | uniq -c | sort -rn | head -40
