#!/bin/bash
# One GPU-box pass (via gpurun): parity tests, smoke, both bench arms, ncu launch list, ncu --set full of the
# library's kernels inside one eager train step.  Everything lands in gpurun_out/.
# usage: scripts/gpu_round.sh [tag] [workload]
TAG=${1:-s5}
WL=${2:-cfg2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu_$TAG.txt 2>&1
(nproc; free -g | head -2) >> $O/gpu_$TAG.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > $O/pytest_gpu_$TAG.log
  tail -8 $O/pytest_gpu_$TAG.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1
  tail -3 $O/smoke_$TAG.log
fi
timeout 600 python bench.py --workload $WL --steps 20 --warmup 5 > $O/bench_${WL}_$TAG.json 2> $O/bench_${WL}_$TAG.err
echo "bench rc=$?"; tail -c 4000 $O/bench_${WL}_$TAG.json; tail -4 $O/bench_${WL}_$TAG.err
if [ -z "$SKIP_REF" ]; then
  timeout 600 python bench.py --impl reference --workload $WL --steps 5 --warmup 3 > $O/bench_ref_${WL}_$TAG.json 2> $O/bench_ref_${WL}_$TAG.err
  echo "ref rc=$?"; tail -c 1500 $O/bench_ref_${WL}_$TAG.json
fi
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_${WL}_$TAG.csv \
   python bench.py --workload $WL --steps 1 --warmup 3 --no-graph --skip-cpu --skip-roofline > $O/ncu_list_${WL}_$TAG.log 2>&1
echo "ncu list rc=$?"; tail -2 $O/ncu_list_${WL}_$TAG.log
if [ -z "$SKIP_FULL" ]; then
  KRE=${KRE:-'regex:tc_pointwise|tc_wgrad_kernel|analysis_fused|modes_gemm|cgemm_kernel|cskinny|coldft|pointwise_small|wgrad_small_partial'}
  timeout 1200 ncu --set full --clock-control none --import-source on -k "$KRE" -c ${NFULL:-90} -f -o $O/full_${WL}_$TAG \
     python bench.py --workload $WL --steps 1 --warmup 3 --no-graph --skip-cpu --skip-roofline > $O/ncu_full_${WL}_$TAG.log 2>&1
  echo "ncu full rc=$?"; tail -2 $O/ncu_full_${WL}_$TAG.log
  ncu -i $O/full_${WL}_$TAG.ncu-rep --page raw --csv > $O/full_${WL}_${TAG}_raw.csv 2>/dev/null
  ls -la $O | tail -20
fi
