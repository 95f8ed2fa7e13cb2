#!/bin/bash
# full GPU validation on an N-GPU box: whole -m gpu suite, then bench lines with the parity gate.  Usage: bash tools/gpu_val.sh <tag> "<N list>" [strong 1|0]
TAG=${1:-val}; NS=${2:-"2"}; STRONG=${3:-1}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest_gpu.log | cut -c1-400
run() { NAME=$1; N=$2; shift 2; ENVS=""; while [ "$1" != "--" ]; do ENVS="$ENVS $1"; shift; done; shift
  if [ "$N" == "1" ]; then env $ENVS timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu "$@" > $OUT/$NAME.json 2> $OUT/$NAME.err
  else env $ENVS timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 "$@" > $OUT/$NAME.json 2> $OUT/$NAME.err; fi
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/$NAME.json").read().strip().splitlines()[-1])
    if "error" in d: print("$NAME ERROR", json.dumps(d)[:1200])
    else:
        s=d.get("sustained",{}); p=d.get("parity",{})
        print("$NAME N=%d %.2f G per-GPU %.2f G | sustained %.2f G | parity fixtures %s bench-size %s | halo %s | %s"%(d["n_gpus"], d["value"]/1e9, d["value"]/1e9/d["n_gpus"], s.get("value",0)/1e9, p.get("fixtures_worst_rel_l2"), p.get("bench_size_rel_l2_max_over_ranks"), d["run"]["halo"][:12], d["config"]["workload"][:64]))
except Exception as ex:
    print("$NAME failed", ex); print(open("$OUT/$NAME.err").read()[-2500:])
PY
}
run bench_n1 1 --
for N in $NS; do
run bench_metis_n$N $N --
run bench_slab_n$N $N -- --partition rcb --shape bar
if [ "$STRONG" == "1" ]; then run bench_strong_n$N $N -- --scaling strong --sustain-s 1; fi
done
if [ "$STRONG" == "1" ]; then run bench_strong_n1 1 -- --scaling strong --sustain-s 1; fi
