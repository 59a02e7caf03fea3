#!/bin/bash
# round 2, GPU call 17: D-infinity links evaluated one per lane; overlapped two-tool host call; cheaper k_deps_dinf
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'sweep\|visits\|passed\|failed\|metric' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-700))"; }
step tests_gpu_v11 900 python -m pytest tests/test_gpu_parity.py -x -q
step stats11_16384 300 python scripts/sweep_stats.py 16384 2
step stats11_65536 600 python scripts/sweep_stats.py 65536 2
TAUDEM_B200_TIMING=1 step stats11_65536_t 600 python scripts/sweep_stats.py 65536 1
step bench11 1500 python bench.py --no-cpu
