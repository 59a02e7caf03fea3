#!/bin/bash
# round 2, GPU call 37: dinfdecayaccum on the GPU, file-level tools after the warm-up / quick-exit change (with the wall-clock trace)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'passed\|failed\|files\|Error\|error' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-2500))"; }
step tests_decay_cli 600 python -m pytest tests/test_gpu_parity.py -x -q -k "decay or file_level_cli or extreme or pointwise"
TAUDEM_B200_TRACE=1 step files_16384_trace 600 python bench.py --files --gpus 1 --size 16384 --no-cpu
