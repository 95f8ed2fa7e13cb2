#!/bin/bash
# A/B of stage-kernel variants on one GPU.  Usage (under gpurun): bash tools/gpu_ab.sh <tag> "<kernel:lib ...>"
#   kernel = value of DGTD_B200_KERNEL ("default" = unset), lib = file under dgtd_b200/ab/ ("tree" = the in-tree library)
TAG=${1:-ab}; VS=${2:-"default:tree"}; ORDERS=${3:-"3"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for V in $VS; do
K=${V%%:*}; L=${V#*:}
if [ "$K" == "default" ]; then unset DGTD_B200_KERNEL; else export DGTD_B200_KERNEL=$K; fi
if [ "$L" == "tree" ]; then unset DGTD_B200_LIB; else export DGTD_B200_LIB=$PWD/dgtd_b200/ab/$L; fi
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q > $OUT/pytest_${K}_$L.log 2>&1; echo "$V pytest exit $?"; tail -1 $OUT/pytest_${K}_$L.log
for P in $ORDERS; do
EXTRA=""; if [ "$P" == "4" ]; then EXTRA="--cubes 26"; fi
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --e2e-steps 1 --order $P $EXTRA > $OUT/bench_${K}_${L}_p$P.json 2> $OUT/bench_${K}_${L}_p$P.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${K}_${L}_p$P.json").read().strip().splitlines()[-1])
    s=d.get("sustained",{})
    print("AB $V p$P: burst %.2f G frac %.3f | sustained %.2f G (%.0f MHz, %.0f W) | %s"%(d["value"]/1e9, d["roofline"]["frac"], s.get("value",0)/1e9, s.get("sm_mhz") or 0, s.get("power_w_max") or 0, d["roofline"]["kernel"][:60]))
except Exception as ex: print("AB $V p$P failed", ex); print(open("$OUT/bench_${K}_${L}_p$P.err").read()[-1500:])
PY
done
done
