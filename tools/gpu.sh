#!/bin/bash
# One parameterised driver for every gpurun call of a round (replaces the
# per-round one-offs).  Usage:
#   gpurun --timeout 1500 [--gpus N] -- 'bash tools/gpu.sh <tag> <stage> [<stage> ...]'
# Output goes to gpurun_out/<tag>/ (copy what should be judged into profiles/<tag>/).
# Stages (N = number of visible GPUs):
#   smoke        __graft_entry__.build() + smoke()
#   tests        pytest tests/ -x -q -m gpu
#   multirank    tests/multirank_parity.py under torchrun, world = N (small,full[,nvls])
#   bench        bench.py --impl reference, then bench.py, as the driver runs them (torchrun if N>1)
#   bench:ARGS   bench.py with extra args (comma separated, e.g. bench:--path,fused,--no-mnist)
#   launches     ncu launch list of the default bench command (1 GPU)
#   ncufull:K    ncu --set full of kernel regex K in the default bench command (1 GPU)
#   nvlink       ncu nvlrx/nvltx byte counters of the cross-GPU kernels, one process driving 2 GPUs
#   sass         cuobjdump -sass extracts (UBLKCP in k_list_tma, multimem in k_round_mc)
#   pcie         tools/bench_pcie.py on all N GPUs at once (the e2e ceiling)
#   p2p          tools/bench_p2p.py (NVLink peaks: uni, duplex, incast, all pairs)
#   nvls         tools/bench_nvls.py (single-process NVLS primitives)
#   sweep:ARGS   tools/bench_sweep.py under torchrun (world = N) with extra args
#   tfrun        tfrun -w 2 -s 1 mnist_replica steady-state step time
TAG=${1:-r00}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
N=$(nvidia-smi -L | wc -l)
nvidia-smi -L > $OUT/gpus.txt 2>&1; nproc >> $OUT/gpus.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
PORT=29500
trun() { PORT=$((PORT+1)); python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $PORT "${@:2}"; }
say() { echo "$@" | tee -a $OUT/summary.txt; }
T=${STAGE_TIMEOUT:-600}   # per-stage limit (seconds): box time is budgeted, a hang must not eat it
BENCH_NCU_ARGS="--steps 3 --warmup 3 --no-mnist --no-cpu-baseline --no-e2e --no-staged --no-verify"
for stage in "$@"; do
  name=${stage%%:*}; arg=""; [[ "$stage" == *:* ]] && arg=${stage#*:}
  extra=${arg//,/ }
  say "== $stage (N=$N)"
  case $name in
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; say "smoke rc=$?"; tail -2 $OUT/smoke.log ;;
    tests)
      ( time timeout $T python -m pytest tests/ -x -q -m gpu ) > $OUT/pytest_gpu.log 2>&1; say "pytest rc=$?"; tail -12 $OUT/pytest_gpu.log | cut -c1-400 ;;
    multirank)
      cases=${arg:-small,full,nvls}
      ( time timeout $T python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29411 tests/multirank_parity.py --cases $cases ) > $OUT/multirank_n$N.log 2>&1; say "multirank rc=$?"
      grep -E "^CASE|MULTIRANK|rank [0-9]+\]" $OUT/multirank_n$N.log | cut -c1-300 | tail -50 ;;
    bench)
      if [ -z "$arg" ]; then
        if [ $N -gt 1 ]; then
          timeout $T bash -c "$(declare -f trun); PORT=29600; trun $N bench.py --impl reference --gpus $N --steps 20 --warmup 5" > $OUT/bench_reference_n$N.json 2> $OUT/bench_reference_n$N.err
          ( time timeout $T bash -c "$(declare -f trun); PORT=29610; trun $N bench.py --gpus $N --steps 20 --warmup 5" ) > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; say "bench rc=$?"
        else
          timeout $T python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference_n1.json 2> $OUT/bench_reference_n1.err
          ( time timeout $T python bench.py --gpus 1 --steps 20 --warmup 5 ) > $OUT/bench_n1.json 2> $OUT/bench_n1.err; say "bench rc=$?"
        fi
        cut -c1-600 $OUT/bench_reference_n$N.json; grep '^{' $OUT/bench_n$N.json | cut -c1-3000; tail -5 $OUT/bench_n$N.err
      else
        f=$OUT/bench_n${N}_$(echo "$arg" | tr -c 'a-zA-Z0-9\n' '_' | cut -c1-60)
        if [ $N -gt 1 ]; then
          ( time timeout $T bash -c "$(declare -f trun); PORT=$((29700 + RANDOM % 200)); trun $N bench.py --gpus $N $extra" ) > $f.json 2> $f.err; say "bench $arg rc=$?"
        else
          ( time timeout $T python bench.py --gpus 1 $extra ) > $f.json 2> $f.err; say "bench $arg rc=$?"
        fi
        grep '^{' $f.json | cut -c1-2500; tail -5 $f.err
      fi ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
        python bench.py $BENCH_NCU_ARGS > $OUT/ncu_launch.log 2>&1; say "ncu list rc=$?" ;;
    ncufull)
      k=${arg:-k_apply}
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o $OUT/prof_$k \
        python bench.py $BENCH_NCU_ARGS > $OUT/ncu_full_$k.log 2>&1; say "ncu full rc=$?"
      ncu -i $OUT/prof_$k.ncu-rep --page raw --csv > $OUT/prof_${k}_raw.csv 2>/dev/null ;;
    nvlink)
      timeout $T ncu --metrics gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_data_user.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none -k regex:'k_apply|k_mc_' --csv --log-file $OUT/nvlink_counters.csv \
        python tools/prof_nvlink.py > $OUT/nvlink_counters.log 2>&1; say "ncu nvlink rc=$?"; tail -3 $OUT/nvlink_counters.log ;;
    sass)
      cuobjdump -sass tfmesos_b200/lib/libpsx.so > $OUT/libpsx.sass 2>&1
      grep -c UBLKCP $OUT/libpsx.sass | sed 's/^/UBLKCP lines: /' | tee -a $OUT/summary.txt ;;
    p2p)
      timeout 900 python tools/bench_p2p.py > $OUT/p2p_n$N.json 2> $OUT/p2p_n$N.err; say "p2p rc=$?"; cut -c1-3000 $OUT/p2p_n$N.json; tail -3 $OUT/p2p_n$N.err ;;
    pcie)
      if [ $N -gt 1 ]; then
        timeout 600 bash -c "$(declare -f trun); PORT=29480; trun $N tools/bench_pcie.py" > $OUT/pcie_n$N.json 2> $OUT/pcie_n$N.err; say "pcie rc=$?"
      else
        timeout 600 python tools/bench_pcie.py > $OUT/pcie_n$N.json 2> $OUT/pcie_n$N.err; say "pcie rc=$?"
      fi
      cat $OUT/pcie_n$N.json; tail -3 $OUT/pcie_n$N.err ;;
    nvls)
      timeout 600 python tools/bench_nvls.py > $OUT/nvls_n$N.json 2> $OUT/nvls_n$N.err; say "nvls rc=$?"; cat $OUT/nvls_n$N.json; tail -3 $OUT/nvls_n$N.err ;;
    sweep)
      f=$OUT/sweep_n${N}_$(echo "$arg" | tr -c 'a-zA-Z0-9\n' '_' | cut -c1-60)
      if [ $N -gt 1 ]; then
        timeout $T bash -c "$(declare -f trun); PORT=$((29900 + RANDOM % 90)); trun $N tools/bench_sweep.py $extra" > $f.jsonl 2> $f.err; say "sweep rc=$?"
      else
        timeout $T python tools/bench_sweep.py $extra > $f.jsonl 2> $f.err; say "sweep rc=$?"
      fi
      cut -c1-260 $f.jsonl | tail -40; tail -3 $f.err ;;
    tfrun)
      timeout $T python tools/bench_tfrun.py > $OUT/tfrun_mnist_replica.json 2> $OUT/tfrun_mnist_replica.err; say "tfrun rc=$?"; cat $OUT/tfrun_mnist_replica.json; tail -5 $OUT/tfrun_mnist_replica.err ;;
    *) say "unknown stage $stage" ;;
  esac
done
ls -la $OUT | tail -40
