#!/bin/bash
# multi-GPU pass: partitioned parity + weak-scaling bench lines.  Usage (gpurun --gpus N): bash tools/gpu_mp.sh <tag> "<N list>"
TAG=${1:-mp}; NS=${2:-"2"}; NCCLNS=${3:-""}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > $OUT/pytest_mp.log 2>&1; echo "pytest exit $?"; tail -30 $OUT/pytest_mp.log | cut -c1-400
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/bench_n1.json 2> $OUT/bench_n1.err
for N in $NS; do
if [[ " $NCCLNS " == *" $N "* ]]; then
DGTD_B200_HALO=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_nccl_n$N.json 2> $OUT/bench_nccl_n$N.err
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
tail -3 $OUT/bench_n$N.err
done
python - <<PY
import json,glob
b=None
for f in sorted(glob.glob("$OUT/bench_n?.json")+glob.glob("$OUT/bench_nccl*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, "no line", ex); continue
    if d["n_gpus"]==1: b=d["value"]
    print(f, "N=%d %.2f G  per-GPU %.2f G  eff %.3f  launches %d"%(d["n_gpus"], d["value"]/1e9, d["value"]/1e9/d["n_gpus"], d["value"]/d["n_gpus"]/b if b else 0, d["gpu_launches"]))
PY
