#!/bin/bash
# 1-GPU call: new tests + PCIe ceilings + bench.
TAG=${1:-r10}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pcie" | tee -a $OUT/summary.txt
timeout 200 python tools/bench_pcie.py > $OUT/pcie.json 2> $OUT/pcie.err; echo "pcie rc=$?" | tee -a $OUT/summary.txt; cat $OUT/pcie.json; tail -2 $OUT/pcie.err
echo "== pytest new" | tee -a $OUT/summary.txt
timeout 1200 python -m pytest tests/test_gpu_examples.py tests/test_gpu_checkpoint.py tests/test_gpu_parity.py -m gpu -q -x -k "to_aggregate or endpoints or 2_pow_31 or batched or resume" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -8 $OUT/pytest.log | cut -c1-300
echo "== bench" | tee -a $OUT/summary.txt
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench.json | cut -c1-2500; tail -4 $OUT/bench.err
echo "== ops" | tee -a $OUT/summary.txt
timeout 200 python tools/bench_ops.py > $OUT/ops.json 2>> $OUT/ops.err; cat $OUT/ops.json
echo "== sweep N=1 small sizes" | tee -a $OUT/summary.txt
SWEEP_MAX=$((16<<20)) timeout 300 python tools/bench_sweep.py > $OUT/sweep_n1.jsonl 2> $OUT/sweep.err; grep '^{' $OUT/sweep_n1.jsonl | cut -c1-200
