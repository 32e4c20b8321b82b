#!/bin/bash
# 2-GPU call: P2P peak, sweep N=2, bf16 tests, bench N=2 bf16 (ResNet bucket, async).
TAG=${1:-r05}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== p2p" | tee -a $OUT/summary.txt
timeout 300 python tools/bench_p2p.py > $OUT/p2p.json 2> $OUT/p2p.err; echo "p2p rc=$?" | tee -a $OUT/summary.txt
cat $OUT/p2p.json; tail -3 $OUT/p2p.err
echo "== pytest bf16 + nmf example" | tee -a $OUT/summary.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_examples.py -m gpu -q -k "bf16 or matrix_factorization or fused" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -5 $OUT/pytest.log | cut -c1-300
echo "== sweep N=2" | tee -a $OUT/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/bench_sweep.py > $OUT/sweep_n2.jsonl 2> $OUT/sweep_n2.err; echo "sweep n2 rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/sweep_n2.jsonl | cut -c1-200; tail -3 $OUT/sweep_n2.err
echo "== sweep N=1" | tee -a $OUT/summary.txt
timeout 600 python tools/bench_sweep.py > $OUT/sweep_n1.jsonl 2> $OUT/sweep_n1.err; echo "sweep n1 rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/sweep_n1.jsonl | cut -c1-200; tail -3 $OUT/sweep_n1.err
echo "== bench N=2 resnet50 bucket, bf16 wire, async" | tee -a $OUT/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --workload resnet50_bucket --wire bf16 --mode async --no-mnist --no-cpu-baseline > $OUT/bench_n2_resnet_bf16.json 2> $OUT/bench_n2_resnet_bf16.err; echo "rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench_n2_resnet_bf16.json; tail -3 $OUT/bench_n2_resnet_bf16.err
ls -la $OUT
