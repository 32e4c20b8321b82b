#!/bin/bash
# 1-GPU call: full gpu suite, bench (fused default + staged + e2e + mnist), ncu traffic capture.
TAG=${1:-r03}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu" | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -12 $OUT/pytest_gpu.log
echo "== bench" | tee -a $OUT/summary.txt
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench.json; tail -5 $OUT/bench.err
echo "== ncu launch list (default bench, fused)" | tee -a $OUT/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-mnist --no-cpu-baseline --no-e2e --no-staged > $OUT/ncu_launch.log 2>&1; echo "ncu list rc=$?" | tee -a $OUT/summary.txt
echo "== ncu full: fused round kernel on the nmf_scaled W shard" | tee -a $OUT/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apply -s 6 -c 1 -o $OUT/prof_round_nmf \
    python bench.py --steps 3 --warmup 3 --no-mnist --no-cpu-baseline --no-e2e --no-staged > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $OUT/summary.txt
echo "== list push/pull micro-bench (TMA vs ld/st vs per-variable)" | tee -a $OUT/summary.txt
timeout 300 python tools/bench_lists.py > $OUT/bench_lists.json 2> $OUT/bench_lists.err; echo "lists rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_lists.json; tail -3 $OUT/bench_lists.err
ls -la $OUT
