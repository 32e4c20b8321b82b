#!/bin/bash
# 4-GPU single-process call: NVLS primitives at N=4.  Short.
TAG=${1:-r19}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python tools/bench_nvls.py > $OUT/nvls_n4.json 2> $OUT/nvls_n4.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/nvls_n4.json; tail -5 $OUT/nvls_n4.err
