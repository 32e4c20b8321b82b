#!/bin/bash
# 2-GPU call: NVLS primitives test + micro-bench.  Short.
TAG=${1:-r18}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest nvls" | tee -a $OUT/summary.txt
timeout 300 python -m pytest tests/test_gpu_nvls.py -m gpu -q -rs > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -15 $OUT/pytest.log | cut -c1-600
echo "== bench nvls" | tee -a $OUT/summary.txt
timeout 300 python tools/bench_nvls.py > $OUT/nvls.json 2> $OUT/nvls.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/nvls.json; tail -5 $OUT/nvls.err
