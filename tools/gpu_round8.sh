#!/bin/bash
# 8-GPU call (charged 8x): N=8 bench with the counter rendez-vous.  Short.
TAG=${1:-r08}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== bench N=8" | tee -a $OUT/summary.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "bench n8 rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench_n8.json; tail -12 $OUT/bench_n8.err
ls -la $OUT
