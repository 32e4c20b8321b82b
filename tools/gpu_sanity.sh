#!/bin/bash
# 1-GPU sanity of the final tree: smoke, gpu suite, both bench arms.
TAG=${1:-r20}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt; tail -1 $OUT/smoke.log
timeout 900 python -m pytest tests/ -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt; tail -3 $OUT/pytest_gpu.log | cut -c1-200
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 > $OUT/bench_reference.json 2>/dev/null; echo "ref rc=$?" | tee -a $OUT/summary.txt
timeout 300 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench.json | cut -c1-300; tail -3 $OUT/bench.err
