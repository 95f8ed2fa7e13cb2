#!/bin/bash
# 1-GPU validation: full GPU test suite, smoke, then the N=1 bench lines with the face-sharing group order (default) against
# plain Morton order (DGTD_B200_ORDER=morton).  Usage: bash tools/gpu_val2.sh <tag>
TAG=${1:-val2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
line() { python -c "import json;d=json.loads(open('$OUT/$1.json').read().strip().splitlines()[-1]);s=(d.get('sustained') or {}).get('value');print('$1: %.2f G frac %.3f sustained %s'%(d['value']/1e9, d['roofline']['frac'], s and round(s/1e9,1)))" 2>/dev/null || { echo "$1 failed"; tail -5 $OUT/$1.err; }; }
for ORD in grow morton; do
  if [ "$ORD" == "morton" ]; then export DGTD_B200_ORDER=morton; else unset DGTD_B200_ORDER; fi
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu --sustain-s 2 --e2e-steps 1 > $OUT/c5_p3_$ORD.json 2> $OUT/c5_p3_$ORD.err; line c5_p3_$ORD
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu --sustain-s 0 --e2e-steps 1 --order 4 --cubes 26 > $OUT/c5_p4_$ORD.json 2> $OUT/c5_p4_$ORD.err; line c5_p4_$ORD
  if [ "$ORD" == "grow" ]; then timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu --sustain-s 0 --e2e-steps 1 --order 2 > $OUT/c5_p2_$ORD.json 2> $OUT/c5_p2_$ORD.err; line c5_p2_$ORD; fi
  timeout 300 python bench.py --workload c4 --steps 40 --warmup 5 --no-cpu --sustain-s 1 --e2e-steps 1 > $OUT/c4_$ORD.json 2> $OUT/c4_$ORD.err; line c4_$ORD
  if [ "$ORD" == "grow" ]; then timeout 300 python bench.py --workload c3 --steps 40 --warmup 5 --no-cpu --sustain-s 0 --e2e-steps 1 > $OUT/c3_$ORD.json 2> $OUT/c3_$ORD.err; line c3_$ORD; fi
done
unset DGTD_B200_ORDER
echo "elapsed $SECONDS s"
