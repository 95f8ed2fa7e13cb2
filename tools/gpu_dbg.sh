#!/bin/bash
# quick multi-GPU check with the bench-size parity.  Usage (gpurun --gpus N): bash tools/gpu_dbg.sh <tag> [N]
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
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -1 $OUT/pytest.log
for V in tree static; do
if [ "$V" == "tree" ]; then unset DGTD_B200_LIB; else export DGTD_B200_LIB=$PWD/dgtd_b200/ab/lib_$V.so; fi
for P in 3 4; do EXTRA=""; if [ "$P" == "4" ]; then EXTRA="--cubes 26"; fi
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --sustain-s 0 --e2e-steps 1 --order $P $EXTRA > $OUT/n1_${V}_p$P.json 2> $OUT/n1_${V}_p$P.err; python -c "import json;d=json.loads(open('$OUT/n1_${V}_p$P.json').read().strip().splitlines()[-1]);print('n1 $V p$P: %.2f G'%(d['value']/1e9))"
done; done
unset DGTD_B200_LIB
run w32_metis --
run w32_slab -- --partition rcb --shape bar
run s64_metis -- --scaling strong --cubes 64
