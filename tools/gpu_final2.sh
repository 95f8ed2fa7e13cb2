#!/bin/bash
# final 2-GPU check: world-2 multirank parity tests (both halo paths, lagging-rank stress) and the bench line as the driver launches it
OUT=gpurun_out/${1:-final2}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q > $OUT/pytest_multirank.log 2>&1; echo "multirank pytest exit $?"; tail -2 $OUT/pytest_multirank.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench exit $?"; tail -c 600 $OUT/bench_n2.json | head -c 600; echo
python -c "import json;d=json.loads(open('$OUT/bench_n2.json').read().strip().splitlines()[-1]);print('n2: %.1f G parity %s halo %s'%(d['value']/1e9, d['config'].get('parity'), d['run']['halo']))"
echo "elapsed $SECONDS s"
