#!/bin/bash
# One gpurun call that refreshes the measured evidence of a round (run from the repo root on the GPU box):
#   gpurun --timeout 1500 -- 'bash tools/gpu_evidence.sh [tests] [bench] [launches] [ncu]'
# Everything lands in gpurun_out/; summaries are then copied into profiles/ by hand (tools/ncu_summary.py).
set -u
mkdir -p gpurun_out
what="${*:-tests bench launches ncu}"
BENCH_SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
for w in $what; do
  case $w in
    tests)
      timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
      echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log ;;
    bench)
      python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
      echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_n1.json ;;
    benchshort)
      python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_n1_short.json 2> gpurun_out/bench_n1_short.err
      echo "bench rc=$?"; cut -c1-900 gpurun_out/bench_n1_short.json ;;
    launches)
      $BENCH_SHORT > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
        --log-file gpurun_out/launches.csv $BENCH_SHORT > gpurun_out/ncu_launches.log 2>&1
      echo "launch list rc=$?" ;;
    ncu)
      # the five biggest kernels of a step, one launch each, after the warm-up steps (5 kernels x 4 passes skipped)
      ncu --set full --clock-control none --import-source on -k 'regex:k_pairs_l1_imma2|k_verify_unite|k_pack_sketch_rows16|k_radix_sort|k_pairs_l2_unit' \
        --launch-skip 20 --launch-count 5 -f -o gpurun_out/top5 $BENCH_SHORT > gpurun_out/ncu_top5.log 2>&1
      echo "ncu rc=$?" ;;
  esac
done
