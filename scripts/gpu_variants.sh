#!/bin/bash
# A/B on the B200 box: Jacobian pass time for library variants (goal_b200/libgoal_b200_<v>.so, built with
# make EXTRA=... OUT=...) x option sets.  Usage: scripts/gpu_variants.sh tag "v1 v2 .." "opts1 opts2 .."   (opts: k=v[,k=v] or -)
set -u
TAG=$1; VARS=$2; OPTS=$3
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/${TAG}_variants.jsonl
for v in $VARS; do
  lib=goal_b200/libgoal_b200.so; [ "$v" != "base" ] && lib=goal_b200/libgoal_b200_$v.so
  for o in $OPTS; do
    [ "$o" = "-" ] && o=""
    echo "== $v [$o]"
    GOAL_B200_LIB=$PWD/$lib GX_OPTS=$o timeout 300 python scripts/time_passes.py 128 J2 jac 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"$v\", \"opts\": \"$o\", /" | tee -a $OUT/${TAG}_variants.jsonl | cut -c1-400
  done
done
