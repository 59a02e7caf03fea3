#!/bin/bash
# round 2, GPU call 41: k_deps_d8 on the tile ring, slopearea exponent fast paths: GPU tests that touch them, timings, ncu
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'passed\|failed\|Error\|error' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-1500))"; }
step tests_deps 900 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or live_reference or odd_direction or slopearea or row_strip or edge_shapes or large or crossings or extreme or gridnet or outlets"
step stencils_16384b 600 python scripts/stencil_bench.py 16384 7
grep -v "^{" gpurun_out/stencils_16384b.log | head -30
step ncu_deps 900 ncu --set full --clock-control none --import-source on -k regex:"k_deps_d8|k_slopearea" -c 8 -f -o gpurun_out/prof_r02h python scripts/stencil_bench.py 8192 1
