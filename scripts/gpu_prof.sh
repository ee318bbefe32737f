#!/bin/bash
# ncu launch list + full capture of named kernels under a bench option.  Usage: scripts/gpu_prof.sh tag "kernel_regex" [bench args...]
set -u
TAG=$1; RX=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_launches.log 2>&1
tail -1 $OUT/${TAG}_ncu_launches.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s 1 -c 2 -f -o $OUT/${TAG}_prof \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_full.log 2>&1
tail -1 $OUT/${TAG}_ncu_full.log | cut -c1-200
ls -la $OUT | grep ${TAG}
