#!/bin/bash
# round 2, GPU call 45: full GPU suite with the fresh library (geographic row strips), k_deps_dinf edge path on / off at 65536^2 and 16384^2
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'passed\|failed\|Error\|error' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-1500))"; }
step tests_gpu 900 python -m pytest tests -m gpu -x -q
tail -3 gpurun_out/tests_gpu.log
for e in 1 0; do
  TAUDEM_B200_DEPS_EDGE=$e step deps_edge${e}_65536 300 python scripts/stencil_bench.py 65536 5 k_deps_dinf,k_deps_d8
  grep "k_deps" gpurun_out/deps_edge${e}_65536.log | grep -v "^{"
  TAUDEM_B200_DEPS_EDGE=$e step deps_edge${e}_16384 300 python scripts/stencil_bench.py 16384 7 k_deps_dinf
  grep "k_deps" gpurun_out/deps_edge${e}_16384.log | grep -v "^{"
done
