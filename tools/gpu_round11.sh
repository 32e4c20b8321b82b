#!/bin/bash
# 8-GPU call (charged 8x): N=8 and N=4, fused + staged, no e2e.  Short.
TAG=${1:-r11}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for N in 8 4; do
echo "== bench N=$N" | tee -a $OUT/summary.txt
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench n$N rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench_n$N.json | cut -c1-1800; tail -6 $OUT/bench_n$N.err
done
