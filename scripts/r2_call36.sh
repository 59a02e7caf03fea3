#!/bin/bash
# round 2, GPU call 36 (2 GPUs): multi-GPU aread8 / areadinf behind the executables
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'passed\|failed\|files\|Error\|error' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-1500))"; }
step tests_cli_mgpu 600 python -m pytest tests/test_gpu_parity.py -x -q -k "file_level_cli"
step files_16384_n2 900 python bench.py --files --gpus 2 --size 16384
