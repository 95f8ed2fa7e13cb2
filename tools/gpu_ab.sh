#!/bin/bash
# A/B pass of the warp-per-group kernel variants: parity tests (default variant), bench lines per variant, ncu of the default.
# Usage (under gpurun): bash tools/gpu_ab.sh <tag> "<variants, e.g. 1 0>" [ncu]
TAG=${1:-ab}; VARS=${2:-"1 0"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/smi.txt 2>&1; nproc >> $OUT/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
for V in $VARS; do
DGTD_B200_WGV=$V timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/bench_v$V.json 2> $OUT/bench_v$V.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_v$V.json").read().strip().splitlines()[-1])
print("V=$V", "%.2f G"%(d["value"]/1e9), "frac %.3f"%d["roofline"]["frac"], "launch ms %.4f"%d["roofline"]["avg_launch_ms"], d["roofline"]["kernel"][:40])
PY
done
if [ "$3" == "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage_wg_kernel -s 5 -c 1 -o $OUT/stage_wg -f python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
fi
ls -la $OUT
