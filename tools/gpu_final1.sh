#!/bin/bash
# final 1-GPU check: full GPU suite, smoke, the default bench line (with cpu_baseline) as the driver runs it, c3 side line, launch list
OUT=gpurun_out/${1:-final1}; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 400 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench exit $?"
python -c "import json;d=json.loads(open('$OUT/bench_default.json').read().strip().splitlines()[-1]);print('default: %.2f G frac %.3f sustained %.1f e2e %.1f cpu %.1f M (%s cores) launches %s'%(d['value']/1e9, d['roofline']['frac'], d['sustained']['value']/1e9, d['e2e']['value']/1e9, d['cpu_baseline']['value']/1e6, d['cpu_baseline']['cores'], d['gpu_launches']))"
echo "after bench $SECONDS s"
if [ $SECONDS -lt 300 ]; then
  for ORD in grow morton; do
    if [ "$ORD" == "morton" ]; then export DGTD_B200_ORDER=morton; else unset DGTD_B200_ORDER; fi
    timeout 100 python bench.py --workload c3 --steps 40 --warmup 5 --no-cpu --sustain-s 0 --e2e-steps 1 > $OUT/c3_$ORD.json 2> $OUT/c3_$ORD.err
    python -c "import json;d=json.loads(open('$OUT/c3_$ORD.json').read().strip().splitlines()[-1]);print('c3 $ORD: %.2f G'%(d['value']/1e9))"
  done; unset DGTD_B200_ORDER
fi
if [ $SECONDS -lt 330 ]; then
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --sustain-s 0 --e2e-steps 1 > $OUT/ncu_bench.log 2>&1; echo "ncu exit $?"
fi
echo "elapsed $SECONDS s"
