#!/bin/bash
# Time the passes with several builds of the library (goal_b200/libgoal_b200_<tag>.so). Usage: scripts/gpu_variants.sh tag v1 v2 ...
set -u
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for v in default "$@"; do
  lib=$PWD/goal_b200/libgoal_b200_$v.so; [ "$v" = default ] && lib=$PWD/goal_b200/libgoal_b200.so
  echo "== $v"; GOAL_B200_LIB=$lib timeout 150 python scripts/time_passes.py 128 J2 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'variant':'$v', **{k:round(v,3) for k,v in d.items() if k.endswith('_ms')}}))" | tee -a $OUT/${TAG}_variants.jsonl
done
