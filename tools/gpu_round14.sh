#!/bin/bash
# 2-GPU call: the new pinned-worker test, the whole gpu suite (incl. multigpu), bench N=2.
TAG=${1:-r14}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu (2 GPUs: nothing skipped)" | tee -a $OUT/summary.txt
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -10 $OUT/pytest_gpu.log | cut -c1-300
echo "== bench N=2" | tee -a $OUT/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --no-cpu-baseline > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench_n2.json | cut -c1-600; tail -3 $OUT/bench_n2.err
