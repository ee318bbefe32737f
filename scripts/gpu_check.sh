#!/bin/bash
# Run on the B200 box via gpurun: GPU tests, smoke, bench (both arms), ncu launch list + full capture.
# Usage: scripts/gpu_check.sh [tag]   (outputs under gpurun_out/<tag>_*)
set -u
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt; free -g | head -2 >> $OUT/${TAG}_gpu.txt
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 | tee $OUT/${TAG}_bench_reference.json | cut -c1-300
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/${TAG}_bench.json | cut -c1-400
echo "== bench neohookean"; timeout 600 python bench.py --model neohookean --no-cpu-baseline --no-sizes 2>&1 | tail -1 | tee $OUT/${TAG}_bench_neo.json | cut -c1-300
echo "== ncu launches (same command as the bench, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-sizes > $OUT/${TAG}_ncu_launches.log 2>&1
tail -1 $OUT/${TAG}_ncu_launches.log | cut -c1-120
echo "== ncu full (both kernels of the Jacobian pass, then the residual / localisation kernels)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"patch_pair_kernel|elem_record_kernel" -s 2 -c 2 -f -o $OUT/${TAG}_prof \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-sizes --no-passes --no-checks > $OUT/${TAG}_ncu_full.log 2>&1
tail -1 $OUT/${TAG}_ncu_full.log | cut -c1-120
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"elem_residual_block_kernel|node_partial_sum_kernel" -s 2 -c 2 -f -o $OUT/${TAG}_prof_res \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sizes --no-checks > $OUT/${TAG}_ncu_full_res.log 2>&1
tail -1 $OUT/${TAG}_ncu_full_res.log | cut -c1-120
ls -la $OUT | grep ${TAG}
