#!/bin/bash
# Quick iteration on the B200 box: GPU parity tests, short bench, ncu launch list + full capture of the Jacobian kernels.
# Usage: scripts/gpu_iter.sh tag [pytest -k expr]   (outputs under gpurun_out/<tag>_*)
set -u
TAG=${1:-it}; KEXPR=${2:-}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest gpu"
if [ -n "$KEXPR" ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.txt
else timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.txt; fi
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/${TAG}_bench.json
echo "== passes"; timeout 600 python scripts/time_passes.py 128 J2 2>&1 | tail -1 | tee $OUT/${TAG}_passes.json
echo "== ncu full (kernels of the Jacobian pass)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"patch_pair_kernel|elem_record_kernel" -s 2 -c 2 -f -o $OUT/${TAG}_prof \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
tail -1 $OUT/${TAG}_ncu_full.log | cut -c1-160
ls -la $OUT | grep ${TAG}
