#!/bin/bash
# 2-GPU call: multi/list/example tests + bench N=2 (fused default).
TAG=${1:-r04}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest multi + lists + examples" | tee -a $OUT/summary.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_lists.py tests/test_gpu_examples.py -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -40 $OUT/pytest_gpu.log | cut -c1-400
echo "== bench N=2" | tee -a $OUT/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_n2.json; tail -8 $OUT/bench_n2.err
ls -la $OUT
