#!/bin/bash
# Round-2 single-GPU pass: GPU tests, default bench line (burst + sustained + e2e), c3 same-mesh pair (GPU vs the reference's
# CPU algorithm), per-stage DRAM traffic of the four launches of a step, launch list.  Usage (under gpurun): bash tools/gpu_r2a.sh [tag]
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/smi.txt 2>&1
nproc >> $OUT/smi.txt; free -g >> $OUT/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -c 2500 $OUT/bench_default.json; tail -3 $OUT/bench_default.err
for V in r1 l2hints; do
DGTD_B200_LIB=$PWD/dgtd_b200/ab/lib_$V.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --sustain-s 0 --e2e-steps 1 > $OUT/bench_ab_$V.json 2> $OUT/bench_ab_$V.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_ab_$V.json").read().strip().splitlines()[-1]); print("AB $V %.2f G frac %.3f"%(d["value"]/1e9, d["roofline"]["frac"]))
except Exception as ex: print("AB $V failed", ex)
PY
done
timeout 300 python bench.py --steps 20 --warmup 5 --workload c3 --no-cpu > $OUT/bench_c3_gpu.json 2> $OUT/bench_c3_gpu.err; tail -c 1200 $OUT/bench_c3_gpu.json; tail -3 $OUT/bench_c3_gpu.err
timeout 300 python bench.py --steps 20 --warmup 5 --workload c4 --no-cpu > $OUT/bench_c4_gpu.json 2> $OUT/bench_c4_gpu.err; tail -c 1200 $OUT/bench_c4_gpu.json; tail -3 $OUT/bench_c4_gpu.err
timeout 300 python bench.py --steps 20 --warmup 5 --order 4 --cubes 26 --no-cpu > $OUT/bench_p4.json 2> $OUT/bench_p4.err; tail -c 600 $OUT/bench_p4.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:stage_w --launch-skip 12 --launch-count 8 --csv --log-file $OUT/traffic_stages.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --sustain-s 0 > $OUT/ncu_traffic.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --sustain-s 0 > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage_wg_kernel -s 5 -c 1 -o $OUT/stage_top -f python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --sustain-s 0 > $OUT/ncu_full.log 2>&1
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 --workload c3 > $OUT/bench_c3_ref.json 2> $OUT/bench_c3_ref.err; tail -c 1500 $OUT/bench_c3_ref.json
ls -la $OUT
