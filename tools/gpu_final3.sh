#!/bin/bash
# last check of the final tree: GPU suite, smoke, one ncu --set full capture of the four stage launches of a step
OUT=gpurun_out/${1:-final3}; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
if [ $SECONDS -lt 170 ]; then
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:stage_wg -s 12 -c 4 -o $OUT/stage_wg_final -f python bench.py --steps 2 --warmup 3 --no-cpu --sustain-s 0 --e2e-steps 1 > $OUT/ncu.log 2>&1; echo "ncu exit $?"
fi
echo "elapsed $SECONDS s"
