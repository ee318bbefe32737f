#!/bin/bash
# like gpu_variants.sh, but times every pass (scripts/time_passes.py without "jac").  Usage: scripts/gpu_variants_all.sh tag "v1 v2 .." "opts1 .."
set -u
TAG=$1; VARS=$2; OPTS=$3
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/${TAG}_variants.jsonl
for v in $VARS; do
  lib=goal_b200/libgoal_b200.so; [ "$v" != "base" ] && lib=goal_b200/libgoal_b200_$v.so
  for o in $OPTS; do
    [ "$o" = "-" ] && o=""
    GOAL_B200_LIB=$PWD/$lib GX_OPTS=$o timeout 300 python scripts/time_passes.py 128 J2 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'variant':'$v','opts':'$o', **{k:round(v,3) for k,v in d.items() if k.endswith('_ms')}, 'stages': d.get('stages')}))" | tee -a $OUT/${TAG}_variants.jsonl
  done
done
