#!/bin/bash
# Scaling lines of the c5 workload on an N-GPU box, each through the full parity gate (fixtures + bench-size N-rank vs 1-rank).
# Usage (gpurun --gpus N): bash tools/gpu_scale.sh <tag> <N> [tests]     -> gpurun_out/<tag>/*.json
TAG=${1:-scale}; N=${2:-4}; TESTS=${3:-}; OUT=gpurun_out/$TAG; mkdir -p $OUT
PORT=29520
show() { python - <<PY
import json
try:
    d=json.loads(open("$OUT/$1.json").read().strip().splitlines()[-1]); p=d.get("parity") or {}
    s=(d.get("sustained") or {}).get("value"); e=(d.get("e2e") or {}).get("value")
    print("$1: n=%d %.1f G (%.1f per GPU) sustained %s e2e %s parity %s halo %s clocks %s"%(d["n_gpus"], d["value"]/1e9, d["value"]/1e9/d["n_gpus"], s and round(s/1e9,1), e and round(e/1e9,1), p.get("bench_size_rel_l2_max_over_ranks"), str((d.get("run") or {}).get("halo"))[:12], (d.get("clocks") or {}).get("sm_mhz")))
except Exception as ex: print("$1 failed", ex); print(open("$OUT/$1.err").read()[-1500:])
PY
}
run() { NAME=$1; R=$2; shift 2; ENVS=""; while [ "$1" != "--" ]; do ENVS="$ENVS $1"; shift; done; shift
  PORT=$((PORT+1))
  if [ "$R" == "1" ]; then env $ENVS timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu "$@" > $OUT/$NAME.json 2> $OUT/$NAME.err
  else env $ENVS timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $R --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $R --steps 20 --warmup 5 --no-cpu "$@" > $OUT/$NAME.json 2> $OUT/$NAME.err; fi
  show $NAME
}
nvidia-smi topo -m > $OUT/topo.txt 2>&1
if [ "$N" == "4" ]; then run n1 1 --; fi
run weak_metis_n$N $N --
run strong65_metis_n$N $N -- --scaling strong --e2e-steps 1
if [ -n "$TESTS" ]; then
  timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > $OUT/pytest_multirank.log 2>&1; echo "multirank pytest exit $?"; tail -2 $OUT/pytest_multirank.log
fi
if [ $SECONDS -lt ${BUDGET_S:-300} ]; then run weak_slab_n$N $N -- --partition rcb --shape bar --e2e-steps 1; fi
echo "elapsed $SECONDS s"
