#!/bin/bash
# 1-GPU call: what the driver runs at round end (smoke, pytest -m gpu -x, bench both arms)
# + compute-sanitizer on small parity tests + ncu of the TMA list kernel.
TAG=${1:-r09}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== smoke" | tee -a $OUT/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt; tail -2 $OUT/smoke.log
echo "== pytest -m gpu -x" | tee -a $OUT/summary.txt
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -12 $OUT/pytest_gpu.log | cut -c1-300
echo "== bench reference arm" | tee -a $OUT/summary.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_reference.json | cut -c1-600
echo "== bench" | tee -a $OUT/summary.txt
( time timeout 900 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
grep '^{' $OUT/bench.json; tail -6 $OUT/bench.err
echo "== compute-sanitizer memcheck (small parity cases)" | tee -a $OUT/summary.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lists.py -q -m gpu -x -k "subranges or special or max_slots or unaligned or (apply_bit_exact and 7850) or bf16_wire_end" > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/summary.txt
tail -8 $OUT/sanitizer_memcheck.log | cut -c1-300
echo "== compute-sanitizer racecheck (tma list + apply)" | tee -a $OUT/summary.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_lists.py tests/test_gpu_parity.py -q -m gpu -x -k "unaligned_tensor or (apply_bit_exact and 1023)" > $OUT/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $OUT/summary.txt
tail -8 $OUT/sanitizer_racecheck.log | cut -c1-300
echo "== ncu: tensor-list kernels" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_list -c 4 -o $OUT/prof_lists python tools/bench_lists.py > $OUT/ncu_lists.log 2>&1; echo "ncu lists rc=$?" | tee -a $OUT/summary.txt
ls -la $OUT
