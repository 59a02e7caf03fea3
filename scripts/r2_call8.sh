#!/bin/bash
# round 2, GPU call 8: warp sweep v5 (36-column ring, 26 / 16 workers per SM, SWAR count publish, no in-loop marks)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 8 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-1200))"; }
step tests_gpu_v5 900 python -m pytest tests/test_gpu_parity.py -x -q
step stats_16384 300 python scripts/sweep_stats.py 16384 2
TAUDEM_B200_TIMING=1 step stats_16384_t 300 python scripts/sweep_stats.py 16384 1
TAUDEM_B200_POLL=1 step stats_16384_poll 300 python scripts/sweep_stats.py 16384 2
TAUDEM_B200_WORKERS=13 step stats_16384_w13 300 python scripts/sweep_stats.py 16384 2
TAUDEM_B200_WORKERS=8 step stats_16384_w8 300 python scripts/sweep_stats.py 16384 2
step stats_65536 600 python scripts/sweep_stats.py 65536 1
TAUDEM_B200_TIMING=1 step stats_65536_t 600 python scripts/sweep_stats.py 65536 1
TAUDEM_B200_WORKERS=13 step stats_65536_w13 600 python scripts/sweep_stats.py 65536 1
