#!/bin/bash
# round 2, GPU call 11: what limits the sweep's throughput? workers per SM x fences (experiment switches)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'sweep\|visits' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-900))"; }
export TAUDEM_B200_TIMING=1
TAUDEM_B200_WORKERS=8 step x_w8 600 python scripts/sweep_stats.py 65536 1
TAUDEM_B200_WORKERS=13 step x_w13 600 python scripts/sweep_stats.py 65536 1
TAUDEM_B200_WORKERS=4 step x_w4 600 python scripts/sweep_stats.py 65536 1
TAUDEM_B200_EXP=7 step x_exp7 600 python scripts/sweep_stats.py 65536 1
TAUDEM_B200_EXP=15 step x_exp15 600 python scripts/sweep_stats.py 65536 1
TAUDEM_B200_EXP=8 step x_exp8 600 python scripts/sweep_stats.py 65536 1
unset TAUDEM_B200_TIMING
TAUDEM_B200_EXP=15 step x_exp15_nt 600 python scripts/sweep_stats.py 65536 1
