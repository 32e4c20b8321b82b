#!/bin/bash
# 1-GPU call mirroring the driver's round-end sequence on the final code, plus the
# ncu launch list and one full capture of the dominant kernel for the same command.
TAG=${1:-r17}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== build + smoke" | tee -a $OUT/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt; tail -2 $OUT/smoke.log
echo "== pytest tests/ -x -q -m gpu" | tee -a $OUT/summary.txt
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -8 $OUT/pytest_gpu.log | cut -c1-300
echo "== bench --impl reference" | tee -a $OUT/summary.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?" | tee -a $OUT/summary.txt
cut -c1-400 $OUT/bench_reference.json
echo "== bench" | tee -a $OUT/summary.txt
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench.json; tail -6 $OUT/bench.err
echo "== ncu launch list of the default bench" | tee -a $OUT/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-mnist --no-cpu-baseline --no-e2e --no-staged > $OUT/ncu_launch.log 2>&1; echo "ncu list rc=$?" | tee -a $OUT/summary.txt
echo "== ncu full: dominant kernel, same command" | tee -a $OUT/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apply -s 6 -c 1 -o $OUT/prof_round_final \
    python bench.py --steps 3 --warmup 3 --no-mnist --no-cpu-baseline --no-e2e --no-staged > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $OUT/summary.txt
ls -la $OUT
