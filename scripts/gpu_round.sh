#!/bin/bash
# usage (on the GPU box, through gpurun): scripts/gpu_round.sh TAG [tests] [bench] [launches] [full:WORKLOAD:ROWS:KERNEL_REGEX:SKIP ...]
# Each leg writes into gpurun_out/; ncu legs run only after the same command has exited 0 without ncu.
tag=$1; shift
mkdir -p gpurun_out
for leg in "$@"; do
  case $leg in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1
      echo "tests rc=$? $(tail -1 gpurun_out/${tag}_tests.log)";;
    bench)
      timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
      echo "bench rc=$?"; python scripts/bench_digest.py gpurun_out/${tag}_bench.json;;
    launches)
      for wl in C2 C3 C4 C5; do
        timeout 600 python bench.py --workload $wl --configs none --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/${tag}_l_$wl.json 2>&1 &&
        timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_$wl.csv \
          python bench.py --workload $wl --configs none --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/${tag}_ncu_l_$wl.log 2>&1
        echo "launches $wl rc=$?"
      done;;
    full:*)
      IFS=: read -r _ wl rows kre skip <<< "$leg"
      cmd="python bench.py --workload $wl --configs none --rows $rows --steps 1 --warmup 3 --no-cpu --no-e2e --no-parity"
      timeout 600 $cmd > gpurun_out/${tag}_f_$wl.json 2>&1 &&
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$kre --launch-skip ${skip:-3} --launch-count 1 \
        -f -o gpurun_out/${tag}_${wl}_${kre%%_kernel*} $cmd > gpurun_out/${tag}_ncu_f_$wl.log 2>&1
      echo "full $wl $kre rc=$?";;
    fullsum:*)   # as full:, but only the digest travels back (a round of .ncu-rep files overflows gpurun_out's 64 MiB)
      IFS=: read -r _ wl rows kre skip <<< "$leg"
      cmd="python bench.py --workload $wl --configs none --rows $rows --steps 1 --warmup 3 --no-cpu --no-e2e --no-parity"
      rep=/tmp/${tag}_${wl}_${kre%%_kernel*}
      timeout 600 $cmd > gpurun_out/${tag}_f_$wl.json 2>&1 &&
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$kre --launch-skip ${skip:-3} --launch-count 1 \
        -f -o $rep $cmd > gpurun_out/${tag}_ncu_f_$wl.log 2>&1 &&
      python scripts/ncu_summary.py $rep.ncu-rep sm__pipe_tensor_cycles_active.avg l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_pipe_tc_wavefronts > gpurun_out/${tag}_prof_${wl}_full.txt 2>&1
      echo "fullsum $wl $kre rc=$?"; rm -f $rep.ncu-rep;;
  esac
done
