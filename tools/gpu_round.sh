#!/bin/bash
# One gpurun call: smoke, gpu tests, bench, ncu launch list + one full capture.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L > $OUT/gpus.txt 2>&1
nproc >> $OUT/gpus.txt
echo "== smoke" | tee -a $OUT/summary.txt
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/smoke.log
echo "== pytest -m gpu" | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -15 $OUT/pytest_gpu.log
echo "== bench" | tee -a $OUT/summary.txt
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench.json; tail -5 $OUT/bench.err
timeout 300 python bench.py --path fused --no-mnist --no-cpu-baseline > $OUT/bench_fused.json 2> $OUT/bench_fused.err; echo "bench fused rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_fused.json; tail -5 $OUT/bench_fused.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench ref rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_reference.json
echo "== ncu launch list" | tee -a $OUT/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-mnist --no-cpu-baseline --no-e2e > $OUT/ncu_launch.log 2>&1; echo "ncu list rc=$?" | tee -a $OUT/summary.txt
echo "== ncu full (apply kernel, resnet50 bucket)" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_apply -s 4 -c 2 -o $OUT/prof_apply \
    python bench.py --workload resnet50_bucket --steps 3 --warmup 3 --no-mnist --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_copy -s 8 -c 2 -o $OUT/prof_copy \
    python bench.py --workload resnet50_bucket --steps 3 --warmup 3 --no-mnist --no-cpu-baseline --no-e2e > $OUT/ncu_full_copy.log 2>&1; echo "ncu full copy rc=$?" | tee -a $OUT/summary.txt
ls -la $OUT
