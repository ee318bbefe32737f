#!/bin/bash
# One ncu --set full capture of the named kernel(s) of the bench command.  Usage: scripts/gpu_prof1.sh tag "kernel_regex" [skip] [bench args...]
set -u
TAG=$1; RX=$2; SKIP=${3:-2}; shift 3 || shift $#
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c 1 -f -o $OUT/${TAG}_prof \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_full.log 2>&1
tail -1 $OUT/${TAG}_ncu_full.log | cut -c1-200
ls -la $OUT | grep ${TAG}
