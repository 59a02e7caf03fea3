#!/bin/bash
# round 2, GPU call 42: multi-GPU pitremove / d8flowdir / dinfflowdir behind the executables (ranks sharing the device), k_deps_d8 v3 timings
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'passed\|failed\|Error\|error' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-1500))"; }
step tests_mgpu_flow 900 python -m pytest tests/test_gpu_parity.py -x -q -k "multi_gpu"
tail -30 gpurun_out/tests_mgpu_flow.log
step tests_gpu 900 python -m pytest tests -m gpu -x -q
step stencils_16384c 600 python scripts/stencil_bench.py 16384 7
grep -v "^{" gpurun_out/stencils_16384c.log | head -8
