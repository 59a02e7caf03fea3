#!/bin/bash
# round 2, GPU call 44 (2 GPUs): the N > 1 paths after this session's changes — DistTools over NCCL (peer sweeps, partitioned flats, halo cell
# sizes), the executables with one rank per GPU (peer-memory sweep, fill / flats exchanges), a short bench line at N = 2
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'passed\|failed\|identical\|DIFFERENT\|Error\|error' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-1500))"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
TD_BACKEND=nccl step dist2_nccl 400 $TR scripts/dist_check.py 2001 2300
step tests_cli_mgpu2 600 python -m pytest tests/test_gpu_parity.py -x -q -k "multi_gpu or geographic"
step bench_n2 600 $TR bench.py --gpus 2 --size 16384 --steps 3 --warmup 3 --no-cpu
tail -c 1200 gpurun_out/bench_n2.log
