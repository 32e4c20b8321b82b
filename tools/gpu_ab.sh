#!/bin/bash
# 1-GPU A/B of k_apply build variants (register budget / prefetch depth).
TAG=${1:-r13}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for v in lb3 lb4 lb2 lb2pf lb3pf lb3; do
  for path in fused staged; do
    TFMESOS_PSX_LIB=$PWD/tfmesos_b200/lib/variants/libpsx_$v.so timeout 200 python bench.py --path $path --steps 30 --no-mnist --no-cpu-baseline --no-e2e --no-staged > $OUT/ab_${v}_$path.json 2> $OUT/ab_${v}_$path.err
    grep '^{' $OUT/ab_${v}_$path.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$v $path ms/step %.4f kernel %.4f ms frac %.4f'%(d['ms_per_step'], r['avg_launch_ms'], r['frac']))" | tee -a $OUT/summary.txt
  done
done
