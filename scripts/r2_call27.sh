#!/bin/bash
# round 2, GPU call 27: evidence for profiles/: ncu launch list of the bench command (16384^2), ncu --set full of the final sweep kernels
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 3 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-400))"; }
step ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench16384.csv python bench.py --size 16384 --steps 2 --warmup 1 --no-cpu --no-same-config --e2e-steps 1
step ncu_sweep_final 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep_warp|k_deps_dinf|k_fill_init" -s 0 -c 4 -f -o gpurun_out/prof_r02f python scripts/prof_kernels.py 16384
