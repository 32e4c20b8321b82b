#!/bin/bash
# 8-GPU call (charged 8x): keep it short.  bench N=8 and N=4, nothing else.
TAG=${1:-r06}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L > $OUT/gpus.txt 2>&1
echo "== bench N=8" | tee -a $OUT/summary.txt
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "bench n8 rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench_n8.json; tail -15 $OUT/bench_n8.err
echo "== bench N=4" | tee -a $OUT/summary.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_n4.json 2> $OUT/bench_n4.err; echo "bench n4 rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench_n4.json; tail -8 $OUT/bench_n4.err
ls -la $OUT
