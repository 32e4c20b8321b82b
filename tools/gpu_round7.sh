#!/bin/bash
# 2-GPU call: op-cost micro-bench, many-shards experiment, new tests.
TAG=${1:-r07}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== op costs" | tee -a $OUT/summary.txt
timeout 200 python tools/bench_ops.py > $OUT/ops.json 2> $OUT/ops.err; echo "ops rc=$?" | tee -a $OUT/summary.txt
cat $OUT/ops.json; tail -3 $OUT/ops.err
echo "== pytest new (batch, checkpoint, bf16)" | tee -a $OUT/summary.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_checkpoint.py -m gpu -q -k "batched or resume or restore or bf16_wire_end or launch_counter" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -6 $OUT/pytest.log | cut -c1-300
for S in 2 8 32; do
echo "== bench N=2 fused stripes=$S" | tee -a $OUT/summary.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$((S%10)) bench.py --gpus 2 --stripes $S --steps 10 --no-mnist --no-cpu-baseline --no-e2e --no-staged > $OUT/bench_n2_s$S.json 2> $OUT/bench_n2_s$S.err; echo "rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench_n2_s$S.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step',d['ms_per_step'],'kernel',d['roofline']['avg_launch_ms'],'launches',d['gpu_launches'])"
tail -2 $OUT/bench_n2_s$S.err
done
echo "== bench N=2 staged stripes=8 (overlap push/apply/pull across shards)" | tee -a $OUT/summary.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29559 bench.py --gpus 2 --path staged --stripes 8 --steps 10 --no-mnist --no-cpu-baseline --no-e2e > $OUT/bench_n2_staged_s8.json 2> $OUT/bench_n2_staged_s8.err; echo "rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench_n2_staged_s8.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step',d['ms_per_step'],'value',d['value'])"
echo "== mnist section N=1 with batch path" | tee -a $OUT/summary.txt
timeout 300 python bench.py --workload nmf_reference --steps 5 --no-cpu-baseline --no-e2e --no-staged > $OUT/bench_small.json 2> $OUT/bench_small.err; echo "rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench_small.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('nmf_reference ms/step',d['ms_per_step'],'mnist',d['mnist_replica'])"
ls $OUT
