#!/bin/bash
# 1-GPU call: captured-round test, torchrun-cluster tests, bench (mnist in one graph).
TAG=${1:-r15}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest" | tee -a $OUT/summary.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "captured or torchrun_cluster or fused_round or launch_counter" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -12 $OUT/pytest.log | cut -c1-400
echo "== bench" | tee -a $OUT/summary.txt
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['roofline']['frac'], d['staged_path']['value'], d['e2e']['value'])
print(d['mnist_replica']); print(d['mnist_softmax_sgd'])"
tail -5 $OUT/bench.err
