#!/bin/bash
# First GPU calls after a CPU-only stretch: everything that was only checked on the CPU emulation, each step under its
# own timeout, logs under gpurun_out/.  Groups (one gpurun call each, one GPU unless noted):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_validate_new.sh core'    parity: default path, opt-in schedules, outlets
#   gpurun --timeout 1800 -- 'bash scripts/gpu_validate_new.sh perf'    timings of every schedule, phases, stencils, bench lines
#   gpurun --timeout 1500 -- 'bash scripts/gpu_validate_new.sh extra'   batched flats, row strips on one GPU (gloo)
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_validate_new.sh dist'   strip flats + level sweeps over NCCL
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_validate_new.sh dist8'  peer mode and level sweeps on 8 strips
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local name=$1; shift; echo "=== $name"; ( time timeout "$@" ) > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300))"; }
summary() { grep -h "DIFFERENT\|identical\|passed\|failed\|Error\|error" gpurun_out/*.log | sort | uniq -c | sort -rn | head -40; }

case "${1:-core}" in
dist8)
  # gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_validate_new.sh dist8': peer mode and level sweeps on 8 strips
  TAUDEM_B200_PEER=1 TD_BACKEND=nccl step dist8_tiles_peer 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 \
    --master-addr 127.0.0.1 --master-port 29513 scripts/dist_check.py 6000 5000
  TAUDEM_B200_SWEEP=levels TAUDEM_B200_FLATS=strips TD_BACKEND=nccl step dist8_levels_strips 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 \
    --master-addr 127.0.0.1 --master-port 29514 scripts/dist_check.py 6000 5000
  ;;
dist)
  for mode in "" levels; do
    for flats in "" strips; do
      tag="dist_${mode:-tiles}_${flats:-replicated}"
      TAUDEM_B200_SWEEP=$mode TAUDEM_B200_FLATS=$flats TD_BACKEND=nccl step "$tag" 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 \
        --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py 3000 4100
    done
  done
  ;;
core)
  # the default path (three stencils were rewritten): the parity tests proper
  step tests_gpu 1200 python -m pytest tests -x -q -m gpu
  # the opt-in schedules against the default, bit for bit; outlets against the oracle and the reference executables
  TAUDEM_B200_TEST_EXPERIMENTAL=1 step tests_experimental 600 python -m pytest tests/test_gpu_parity.py -q -k 'experimental or outlets'
  step modes_4096 600 python scripts/sweep_modes.py 4096 tiles,levels:8,levels:24,levels:64,levels:24+river:64,hybrid,walk,walk+river:64 2
  ;;
perf)
  step modes_16384 900 python scripts/sweep_modes.py 16384 tiles,levels:24,levels:48,levels:auto,levels:24+river:64,levels:auto+river:64,hybrid 2
  RIVER_DINF=1 step modes_16384_river_dinf 600 python scripts/sweep_modes.py 16384 levels:24,levels:24+river:64 2
  TAUDEM_B200_TIMING=2 step modes_16384_phases 600 python scripts/sweep_modes.py 16384 levels:64,levels:24+river:64 1
  step perf_16384 600 python scripts/gpu_perf.py 16384
  step bench_16384_levels 900 python bench.py --size 16384 --steps 3 --warmup 3 --no-cpu --sweep levels:24+river:64
  step bench_16384_tiles 900 python bench.py --size 16384 --steps 3 --warmup 3 --no-cpu
  ;;
extra)
  TAUDEM_B200_FLATS_BATCH=64 step tests_flats_batch 600 python -m pytest tests/test_gpu_parity.py -q -k "golden or live_reference or depression"
  TAUDEM_B200_FLATS_BATCH=64 step perf_16384_flats_batch 600 python scripts/gpu_perf.py 16384
  TAUDEM_B200_SWEEP=levels TAUDEM_B200_FLATS=strips TD_BACKEND=gloo step strips_gloo 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 \
    --master-addr 127.0.0.1 --master-port 29512 scripts/dist_check.py 1001 1300
  ;;
esac
summary
