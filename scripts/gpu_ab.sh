#!/bin/bash
# A/B of two builds of the library on the B200 box: GPU tests (new build), then every pass timed with both builds,
# then an ncu launch list of the new build.  Usage: scripts/gpu_ab.sh tag [pytest -k expr]
set -u
TAG=${1:-ab}; KEXPR=${2:-"not beyond_2_31"}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
echo "== passes (new)"; timeout 400 python scripts/time_passes.py 128 J2 2>&1 | tail -1 | tee $OUT/${TAG}_passes_new.json
if [ -f goal_b200/libgoal_b200_base.so ]; then
  echo "== passes (base)"; GOAL_B200_LIB=$PWD/goal_b200/libgoal_b200_base.so timeout 400 python scripts/time_passes.py 128 J2 2>&1 | tail -1 | tee $OUT/${TAG}_passes_base.json
fi
echo "== ncu launch list (new)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
  python scripts/time_passes.py 128 J2 > $OUT/${TAG}_ncu.log 2>&1
python scripts/launch_shares.py $OUT/${TAG}_launches.csv 2>/dev/null | tail -20
