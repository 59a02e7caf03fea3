#!/bin/bash
# round 2, GPU call 14 (2 GPUs): rank-count invariance over NCCL with the sharded scheduler + token termination (peer mode),
# partitioned flats, a 2-GPU bench line
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(grep -h 'identical\|DIFFERENT\|Error\|error\|metric' gpurun_out/$name.log | tr '\n' ' ' | cut -c1-1500))"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
step dist2_peer 600 $TR scripts/dist_check.py 3001 2500
TAUDEM_B200_FLATS=strips step dist2_flats_strips 600 $TR scripts/dist_check.py 2200 1900
TAUDEM_B200_PEER=0 step dist2_rounds 600 $TR scripts/dist_check.py 1500 1300
step bench2_16384 900 $TR bench.py --gpus 2 --size 16384 --steps 3 --warmup 3
