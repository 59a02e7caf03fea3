#!/bin/bash
# round 2, GPU call 38: HEAD after the container re-creation — whole GPU suite, smoke, default bench line
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'passed\|failed\|smoke\|Error\|error' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-2500))"; }
step tests_gpu 900 python -m pytest tests -m gpu -x -q --durations=10
step smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
step bench_default 600 python bench.py
tail -c 1500 gpurun_out/bench_default.log
