#!/bin/bash
# Jacobian pass time under schedule-builder environment settings.  Usage: scripts/gpu_envsweep.sh tag "VAR=a VAR2=b" "VAR=c" ...
set -u
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/${TAG}_envsweep.jsonl
for cfg in "$@"; do
  echo "== [$cfg]"
  env $cfg timeout 300 python scripts/time_passes.py 128 J2 jac 2>&1 | tail -1 | sed "s/^{/{\"env\": \"$cfg\", /" | tee -a $OUT/${TAG}_envsweep.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['env'], d['jacobian_primal_save_ms'])"
done
