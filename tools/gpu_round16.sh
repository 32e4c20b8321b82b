#!/bin/bash
# 4-GPU call (charged 4x): counted rendez-vous at N=4 and N=2, default bench, short.
TAG=${1:-r16}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for N in 4 2; do
echo "== bench N=$N" | tee -a $OUT/summary.txt
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench n$N rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['staged_path']['value'], d['e2e'] and d['e2e']['value'])
print(d['mnist_replica']['global_steps_per_sec'], d['mnist_replica']['worker_compute'], d['mnist_softmax_sgd']['global_steps_per_sec'])"
tail -4 $OUT/bench_n$N.err
done
