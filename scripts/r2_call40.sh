#!/bin/bash
# round 2, GPU call 40: stencil / streaming kernel timings (k_fill_init on the tile ring), ncu --set full of the new kernels
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'passed\|failed\|Error\|error' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-1500))"; }
step stencils_16384 600 python scripts/stencil_bench.py 16384 7
grep -v "^{" gpurun_out/stencils_16384.log | head -30
step stencils_32768 600 python scripts/stencil_bench.py 32768 5
grep -v "^{" gpurun_out/stencils_32768.log | head -30
step ncu_new 900 ncu --set full --clock-control none --import-source on -k regex:"k_fill_init|k_deps_d8|k_threshold|k_slopearea|k_twi" -c 14 -f -o gpurun_out/prof_r02g python scripts/stencil_bench.py 8192 1
ls -la gpurun_out/*.ncu-rep
