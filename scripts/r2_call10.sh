#!/bin/bash
# round 2, GPU call 10: D-infinity shares precomputed by the dependency stencil (8 workers per SM), flat scratch released
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 8 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-1200))"; }
step tests_gpu_v7 900 python -m pytest tests/test_gpu_parity.py -x -q
step stats7_16384 300 python scripts/sweep_stats.py 16384 2
step stats7_65536 600 python scripts/sweep_stats.py 65536 2
TAUDEM_B200_TIMING=1 step stats7_65536_t 600 python scripts/sweep_stats.py 65536 1
step ncu_sweep7 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep_warp" -s 0 -c 2 -f -o gpurun_out/prof_r02d python scripts/prof_kernels.py 16384
