#!/bin/bash
# 2-GPU call: full gpu test suite (incl. multigpu + IPC), bench at N=1 and N=2.
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L > $OUT/gpus.txt 2>&1
nvidia-smi topo -m >> $OUT/gpus.txt 2>&1
echo "== pytest -m gpu" | tee -a $OUT/summary.txt
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -25 $OUT/pytest_gpu.log
echo "== bench N=1" | tee -a $OUT/summary.txt
timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench n1 rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_n1.json; tail -5 $OUT/bench_n1.err
for path in staged fused; do
echo "== bench N=2 $path" | tee -a $OUT/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --path $path --no-cpu-baseline > $OUT/bench_n2_$path.json 2> $OUT/bench_n2_$path.err; echo "bench n2 $path rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_n2_$path.json; tail -8 $OUT/bench_n2_$path.err
done
echo "== bench N=2 stripes=1 (reference-literal placement: W on ps:0=GPU0, H on ps:1=GPU1)" | tee -a $OUT/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --stripes 1 --no-cpu-baseline --no-mnist --no-e2e > $OUT/bench_n2_stripes1.json 2> $OUT/bench_n2_stripes1.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_n2_stripes1.json; tail -5 $OUT/bench_n2_stripes1.err
ls -la $OUT
