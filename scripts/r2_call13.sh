#!/bin/bash
# round 2, GPU call 13: plain stores for interior count words, deliveries first, two fences per visit; warps per SM experiment; ncu
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'sweep\|visits\|passed\|failed' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-900))"; }
step tests_gpu_v9 900 python -m pytest tests/test_gpu_parity.py -x -q
step stats9_16384 300 python scripts/sweep_stats.py 16384 2
step stats9_65536 600 python scripts/sweep_stats.py 65536 2
TAUDEM_B200_TIMING=1 step stats9_65536_t 600 python scripts/sweep_stats.py 65536 1
TAUDEM_B200_TIMING=1 TAUDEM_B200_WORKERS=8 TAUDEM_B200_SMEMPAD=120000 step stats9_65536_8w 600 python scripts/sweep_stats.py 65536 1
TAUDEM_B200_TIMING=1 TAUDEM_B200_WORKERS=13 TAUDEM_B200_SMEMPAD=120000 step stats9_65536_13w 600 python scripts/sweep_stats.py 65536 1
step ncu_sweep9 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep_warp" -s 0 -c 2 -f -o gpurun_out/prof_r02e python scripts/prof_kernels.py 16384
