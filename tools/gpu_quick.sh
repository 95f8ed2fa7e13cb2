#!/bin/bash
# quick A/B of stage kernels: 3-D parity cases + one bench line per kernel.  Usage: bash tools/gpu_quick.sh <tag> "<kernels: default wg wp ws mma generic>"
TAG=${1:-q}; KS=${2:-"default"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for K in $KS; do
if [ "$K" == "default" ]; then unset DGTD_B200_KERNEL; else export DGTD_B200_KERNEL=$K; fi
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_$K.log 2>&1; echo "$K pytest exit $?"; tail -1 $OUT/pytest_$K.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/bench_$K.json 2> $OUT/bench_$K.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_$K.json").read().strip().splitlines()[-1])
print("$K", "%.2f G"%(d["value"]/1e9), "frac %.3f"%d["roofline"]["frac"], "launch ms %.4f"%d["roofline"]["avg_launch_ms"], d["roofline"]["kernel"][:40], d["clocks"])
PY
done
