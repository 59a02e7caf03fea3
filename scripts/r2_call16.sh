#!/bin/bash
# round 2, GPU call 16 (N GPUs): the strong-scaling bench line at 65536^2 with per-tool pipeline timings and the -wg variant
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-4}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 ) > gpurun_out/bench_n${N}_65536.log 2>&1
echo "exit $?"; grep -h "metric\|inputs ready\|Error\|error" gpurun_out/bench_n${N}_65536.log | cut -c1-3000
