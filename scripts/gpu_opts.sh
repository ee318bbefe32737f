#!/bin/bash
# Time the passes under several library option sets. Usage: scripts/gpu_opts.sh tag "k=v,k=v" ...   ("default" = none)
set -u
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for o in "$@"; do
  oo=$o; [ "$o" = default ] && oo=""
  echo "== $o"; GX_OPTS=$oo timeout 150 python scripts/time_passes.py 128 J2 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'opts':'$o', **{k:round(v,3) for k,v in d.items() if k.endswith('_ms')}, 'stages': d.get('stages')}))" | tee -a $OUT/${TAG}_opts.jsonl
done
