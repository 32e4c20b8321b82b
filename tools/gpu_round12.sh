#!/bin/bash
# 2-GPU call: new tests, full default bench at N=2 (incl. e2e with the counter protocol),
# and the literal reference-API path: tfrun + mnist_replica, timed by the script itself.
TAG=${1:-r12}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest (row blocks, multi, checkpoint endpoints)" | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_checkpoint.py -m gpu -q -x -k "row_block or two_gpus or ipc or endpoints" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -6 $OUT/pytest.log | cut -c1-300
echo "== bench N=2 (all sections)" | tee -a $OUT/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench_n2.json | cut -c1-3000; tail -5 $OUT/bench_n2.err
echo "== tfrun mnist_replica, 1 ps + 2 workers (one GPU each), 400 global steps, async" | tee -a $OUT/summary.txt
export PYTHONPATH=$PWD
timeout 600 python script/tfrun -w 2 -s 1 -Gw 1 -- python examples/mnist/mnist_replica.py --ps_hosts {ps_hosts} --worker_hosts {worker_hosts} --job_name {job_name} --worker_index {task_index} --train_steps 400 > $OUT/tfrun_mnist_async.log 2>&1; echo "tfrun rc=$?" | tee -a $OUT/summary.txt
grep -E "Training elapsed|validation cross" $OUT/tfrun_mnist_async.log
echo "== tfrun mnist_replica, 1 ps + 1 worker (README.rst:92), 200 steps" | tee -a $OUT/summary.txt
timeout 600 python script/tfrun -w 1 -s 1 -Gw 1 -- python examples/mnist/mnist_replica.py --ps_hosts {ps_hosts} --worker_hosts {worker_hosts} --job_name {job_name} --worker_index {task_index} --train_steps 200 > $OUT/tfrun_mnist_w1.log 2>&1; echo "tfrun rc=$?" | tee -a $OUT/summary.txt
grep -E "Training elapsed|validation cross" $OUT/tfrun_mnist_w1.log
echo "== embedding_1e6x1e3 workload N=1 (4 GB parameter, the other reading of config #3)" | tee -a $OUT/summary.txt
timeout 600 python bench.py --workload embedding_1e6x1e3 --steps 10 --no-mnist --no-cpu-baseline --no-e2e > $OUT/bench_embedding.json 2> $OUT/bench_embedding.err; echo "rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench_embedding.json | cut -c1-1500; tail -3 $OUT/bench_embedding.err
ls $OUT
