#!/bin/bash
# ncu launch list of one eager train step + ncu --set full of the library's kernels in the first eager step.
# Keeps gpurun_out/ small: the .ncu-rep files are exported to csv pages and deleted unless KEEP_REP=1.
# usage: scripts/gpu_prof.sh [tag] [workload]
TAG=${1:-s5}
WL=${2:-cfg2}
O=gpurun_out
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_${WL}_$TAG.csv \
   python bench.py --workload $WL --steps 1 --warmup 3 --no-graph --skip-cpu --skip-roofline > $O/ncu_list_${WL}_$TAG.log 2>&1
echo "ncu list rc=$?"
KRE=${KRE:-'regex:tc_pointwise|tc_wgrad_kernel|analysis_fused|modes_gemm|cgemm_kernel|cskinny|coldft|pointwise_small|wgrad_small_partial'}
timeout 900 ncu --set full --clock-control none --import-source on -k "$KRE" -c ${NFULL:-48} -f -o $O/full_${WL}_$TAG \
   python bench.py --workload $WL --steps 1 --warmup 3 --no-graph --skip-cpu --skip-roofline > $O/ncu_full_${WL}_$TAG.log 2>&1
echo "ncu full rc=$?"; tail -2 $O/ncu_full_${WL}_$TAG.log
ncu -i $O/full_${WL}_$TAG.ncu-rep --page raw --csv > $O/full_${WL}_${TAG}_raw.csv 2>/dev/null
ncu -i $O/full_${WL}_$TAG.ncu-rep --page source --csv -k regex:tc_pointwise -c 4 > $O/full_${WL}_${TAG}_src_tcpw.csv 2>/dev/null
ncu -i $O/full_${WL}_$TAG.ncu-rep --page source --csv -k regex:tc_wgrad_kernel -c 1 > $O/full_${WL}_${TAG}_src_tcwg.csv 2>/dev/null
ncu -i $O/full_${WL}_$TAG.ncu-rep --page source --csv -k regex:analysis_fused -c 1 > $O/full_${WL}_${TAG}_src_analysis.csv 2>/dev/null
ncu -i $O/full_${WL}_$TAG.ncu-rep --page source --csv -k regex:modes_gemm2 -c 1 > $O/full_${WL}_${TAG}_src_modes.csv 2>/dev/null
ncu -i $O/full_${WL}_$TAG.ncu-rep --page source --csv -k regex:coldft_inv -c 1 > $O/full_${WL}_${TAG}_src_coldft.csv 2>/dev/null
[ -z "$KEEP_REP" ] && rm -f $O/full_${WL}_$TAG.ncu-rep
du -sh $O; ls -la $O | tail -15
