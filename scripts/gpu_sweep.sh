#!/bin/bash
# Quick A/B of kernel options on the B200 box. Usage: scripts/gpu_sweep.sh tag "opt1 opt2 ..." (each opt: k=v[,k=v])
set -u
TAG=${1:-sweep}; shift
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/${TAG}_pytest.txt
for o in "$@"; do
  args=""; for kv in ${o//,/ }; do [ "$kv" != "default" ] && args="$args --opt $kv"; done
  echo "== bench $o"
  timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline $args 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'opt':'$o','value':d['value'],'kernel_ms':d['roofline']['kernel_ms_per_pass'],'frac':d['roofline']['frac'],'clocks':d['clocks']}))" | tee -a $OUT/${TAG}_sweep.jsonl
done
