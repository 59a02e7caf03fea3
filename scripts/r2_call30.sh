#!/bin/bash
# round 2, GPU call 30 (2 GPUs): edge-seeded fill rounds; rank-count invariance again; the 2-GPU bench line at 65536^2
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'identical\|DIFFERENT\|Error\|error\|metric\|inputs ready' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-2500))"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
step dist2b_peer 600 $TR scripts/dist_check.py 3001 2500
step bench_n2_65536 900 $TR bench.py --gpus 2 --steps 3 --warmup 3
