#!/bin/bash
# A/B pass of one kernel variant: parity tests and a bench line.  Usage: bash tools/gpu_wg.sh <tag> <kernel> [ncu]
TAG=${1:-wg}; K=${2:-wg}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export DGTD_B200_KERNEL=$K
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -15 $OUT/pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; tail -c 900 $OUT/bench.json; tail -5 $OUT/bench.err
if [ "$3" == "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage_${K}_kernel -s 5 -c 1 -o $OUT/stage_$K -f python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
fi
