#!/bin/bash
# quick A/B of kernel variants: parity of the 3-D cases per variant + bench line.  Usage: bash tools/gpu_quick.sh <tag> "<variants>"
TAG=${1:-q}; VARS=${2:-"1"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for V in $VARS; do
DGTD_B200_WGV=$V timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_v$V.log 2>&1; echo "V=$V pytest exit $?"; tail -1 $OUT/pytest_v$V.log
DGTD_B200_WGV=$V timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/bench_v$V.json 2> $OUT/bench_v$V.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_v$V.json").read().strip().splitlines()[-1])
print("V=$V", "%.2f G"%(d["value"]/1e9), "frac %.3f"%d["roofline"]["frac"], "launch ms %.4f"%d["roofline"]["avg_launch_ms"], d["roofline"]["kernel"][:40])
PY
done
