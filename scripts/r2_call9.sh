#!/bin/bash
# round 2, GPU call 9: sharded ticket queues (64 shards, done/tail termination) on top of warp sweep v5
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 8 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-1200))"; }
step tests_gpu_v6 900 python -m pytest tests/test_gpu_parity.py -x -q
step stats6_16384 300 python scripts/sweep_stats.py 16384 2
step stats6_65536 600 python scripts/sweep_stats.py 65536 2
TAUDEM_B200_TIMING=1 step stats6_65536_t 600 python scripts/sweep_stats.py 65536 1
TAUDEM_B200_WORKERS=13 step stats6_65536_w13 600 python scripts/sweep_stats.py 65536 1
TAUDEM_B200_WORKERS=8 step stats6_65536_w8 600 python scripts/sweep_stats.py 65536 1
TAUDEM_B200_POLL=1 step stats6_65536_poll 600 python scripts/sweep_stats.py 65536 1
