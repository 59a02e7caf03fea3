#!/bin/bash
# round 2, GPU call 29: validation of the current head: GPU tests, the bench line, smoke
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'passed\|failed\|metric\|smoke' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-400))"; }
step tests_gpu_v13 900 python -m pytest tests -x -q -m gpu
step smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
step bench13 1500 python bench.py
