#!/bin/bash
# A/B of halo-protocol timing variants on N GPUs (bench-size parity reported, not gated).  Usage: bash tools/gpu_ab2.sh <tag> <N> <variant...>
TAG=${1:-ab2}; N=${2:-2}; shift 2; OUT=gpurun_out/$TAG; mkdir -p $OUT
PORT=29540
for V in "$@"; do
  unset DGTD_B200_ORDER
  if [ "$V" == "tree" ]; then unset DGTD_B200_LIB; elif [ "$V" == "MORTON" ]; then unset DGTD_B200_LIB; export DGTD_B200_ORDER=morton; else export DGTD_B200_LIB=$PWD/dgtd_b200/ab/lib_$V.so; fi
  for W in ${WORKLOADS:-weak}; do
    PORT=$((PORT+1)); EXTRA=""; if [ "$W" == "strong" ]; then EXTRA="--scaling strong --cubes 64"; fi
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 40 --warmup 5 --no-cpu --no-gate --parity-report-only --sustain-s 0 --e2e-steps 1 $EXTRA > $OUT/${V}_$W.json 2> $OUT/${V}_$W.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/${V}_$W.json").read().strip().splitlines()[-1]); p=d.get("parity") or {}
    print("$V $W: %.1f G (%.2f per GPU) parity %s differing %s"%(d["value"]/1e9, d["value"]/1e9/d["n_gpus"], p.get("bench_size_rel_l2_max_over_ranks"), p.get("elements_differing")))
except Exception as ex: print("$V $W failed", ex); print(open("$OUT/${V}_$W.err").read()[-800:])
PY
  done
done
unset DGTD_B200_LIB DGTD_B200_ORDER
if [ -z "$SKIP_N1" ]; then timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu --sustain-s 0 --e2e-steps 1 > $OUT/n1.json 2> $OUT/n1.err; python -c "import json;d=json.loads(open('$OUT/n1.json').read().strip().splitlines()[-1]);print('n1: %.2f G'%(d['value']/1e9))"; fi
echo "elapsed $SECONDS s"
