#!/bin/bash
# A/B of the halo hand-shake variants with the bench-size parity check.  Usage (gpurun --gpus N): bash tools/gpu_dbg.sh <tag> [N]
TAG=${1:-dbg}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { NAME=$1; shift; ENVS=""; while [ "$1" != "--" ]; do ENVS="$ENVS $1"; shift; done; shift
  env $ENVS timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-gate --parity-report-only --sustain-s 0 --e2e-steps 1 "$@" > $OUT/$NAME.json 2> $OUT/$NAME.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/$NAME.json").read().strip().splitlines()[-1]); p=d.get("parity",{})
    print("$NAME: %.1f G (%.1f per GPU) parity max %s differing %s of %s halo %s"%(d["value"]/1e9, d["value"]/1e9/d["n_gpus"], p.get("bench_size_rel_l2_max_over_ranks"), p.get("elements_differing"), p.get("elements"), d["run"]["halo"][:10]))
except Exception as ex: print("$NAME failed", ex); print(open("$OUT/$NAME.err").read()[-1500:])
PY
}
L=$PWD/dgtd_b200/ab
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --sustain-s 0 --e2e-steps 1 > $OUT/n1.json 2> $OUT/n1.err; python -c "import json;d=json.loads(open('$OUT/n1.json').read().strip().splitlines()[-1]);print('n1: %.1f G'%(d['value']/1e9))"
run w32_deferred --
run w32_deferred_slab -- --partition rcb --shape bar
run w32_immediate DGTD_B200_LIB=$L/lib_immediate.so --
run w32_noacq DGTD_B200_LIB=$L/lib_noacq.so --
run w32_noacq_slab DGTD_B200_LIB=$L/lib_noacq.so -- --partition rcb --shape bar
run s64_deferred -- --scaling strong --cubes 64
run s64_noacq DGTD_B200_LIB=$L/lib_noacq.so -- --scaling strong --cubes 64
run w32_nccl DGTD_B200_HALO=nccl --
