#!/bin/bash
# debugging pass for the bench-size multi-GPU parity: fence variants / processing order at the failing size.  Usage (gpurun --gpus 2)
TAG=${1:-dbg}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { NAME=$1; shift; ENVS=""; while [ "$1" != "--" ]; do ENVS="$ENVS $1"; shift; done; shift
  env $ENVS timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-gate --parity-report-only --sustain-s 0 --e2e-steps 1 "$@" > $OUT/$NAME.json 2> $OUT/$NAME.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/$NAME.json").read().strip().splitlines()[-1]); p=d.get("parity",{})
    print("$NAME: %.1f G parity max %s global %s halo %s"%(d["value"]/1e9, p.get("bench_size_rel_l2_max_over_ranks"), p.get("bench_size_rel_l2_global"), d["run"]["halo"][:10]))
except Exception as ex: print("$NAME failed", ex); print(open("$OUT/$NAME.err").read()[-1500:])
PY
}
run s64_allfence -- --scaling strong --cubes 64
run s64_allfence_natural DGTD_B200_P2P_ORDER=natural -- --scaling strong --cubes 64
run s64_elected DGTD_B200_LIB=$PWD/dgtd_b200/ab/lib_elected.so -- --scaling strong --cubes 64
run s64_elected_natural DGTD_B200_LIB=$PWD/dgtd_b200/ab/lib_elected.so DGTD_B200_P2P_ORDER=natural -- --scaling strong --cubes 64
run s64_r1lib DGTD_B200_LIB=$PWD/dgtd_b200/ab/lib_r1.so -- --scaling strong --cubes 64
run w32_allfence -- 
run w32_allfence_natural DGTD_B200_P2P_ORDER=natural --
