#!/bin/bash
make -C "$(dirname "$0")/../taudem_b200/csrc" -j8 -s || exit 1     # never ship a stale library
# gpurun with retries on "no box free" (exit 3): scripts/gpurun_retry.sh <timeout> '<command>' [--gpus N]
T=$1; CMD=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
