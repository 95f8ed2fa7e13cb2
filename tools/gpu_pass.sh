#!/bin/bash
# One GPU-box pass: parity tests, bench (default kernel and the A/B kernels), ncu launch list + one full capture.
# Usage (under gpurun): bash tools/gpu_pass.sh [tag] [kernel-regex for the full capture]
TAG=${1:-pass}
KRE=${2:-stage_wg_kernel}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/smi.txt 2>&1
nproc >> $OUT/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -c 1500 $OUT/bench_default.json
for K in ws mma generic; do
DGTD_B200_KERNEL=$K timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/bench_$K.json 2> $OUT/bench_$K.err; tail -c 600 $OUT/bench_$K.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 5 -c 1 -o $OUT/stage_top -f python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/ncu_full.log 2>&1
ls -la $OUT
