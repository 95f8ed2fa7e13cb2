#!/bin/bash
# Round-2 multi-GPU pass.  Usage (gpurun --gpus N): bash tools/gpu_mp2.sh <tag> "<N list>" [tests: 1|0] [strong: 1|0]
TAG=${1:-mp}; NS=${2:-"2"}; TESTS=${3:-1}; STRONG=${4:-1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L | wc -l
run() { # run <name> <N> <env...> -- <bench args>
  NAME=$1; N=$2; shift 2; ENVS=""; while [ "$1" != "--" ]; do ENVS="$ENVS $1"; shift; done; shift
  if [ "$N" == "1" ]; then env $ENVS timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu "$@" > $OUT/$NAME.json 2> $OUT/$NAME.err
  else env $ENVS timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 "$@" > $OUT/$NAME.json 2> $OUT/$NAME.err; fi
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/$NAME.json").read().strip().splitlines()[-1])
    if "error" in d: print("$NAME ERROR", json.dumps(d)[:1500])
    else:
        s=d.get("sustained",{}); p=d.get("parity",{})
        print("$NAME N=%d %.2f G per-GPU %.2f G | sustained %.2f G | parity fixtures %s bench-size %s | halo %s | %s"%(d["n_gpus"], d["value"]/1e9, d["value"]/1e9/d["n_gpus"], s.get("value",0)/1e9, p.get("fixtures_worst_rel_l2"), p.get("bench_size_rel_l2_max_over_ranks"), d["run"]["halo"][:12], d["config"]["workload"][:70]))
except Exception as ex:
    print("$NAME failed", ex); print(open("$OUT/$NAME.err").read()[-2500:])
PY
}
if [ "$TESTS" == "1" ]; then
timeout 1200 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_mfem_shell.py -m gpu -x -q > $OUT/pytest_mp.log 2>&1; echo "pytest exit $?"; tail -30 $OUT/pytest_mp.log | cut -c1-600
fi
run bench_n1 1 --
for N in $NS; do
run bench_metis_n$N $N --
run bench_slab_n$N $N -- --partition rcb --shape bar
run bench_rcbcube_n$N $N -- --partition rcb
run bench_nccl_n$N $N DGTD_B200_HALO=nccl -- --sustain-s 0 --no-gate --parity-steps 0
if [ "$STRONG" == "1" ]; then run bench_strong_n$N $N -- --scaling strong --sustain-s 1; fi
done
if [ "$STRONG" == "1" ]; then run bench_strong_n1 1 -- --scaling strong --sustain-s 1; fi
ls $OUT
