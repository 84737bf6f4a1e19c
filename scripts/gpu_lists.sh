#!/bin/bash
# ncu launch lists (one eager step) for the workloads named on the command line
O=gpurun_out; mkdir -p $O
for wl in "$@"; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_${wl}_s5.csv \
     python bench.py --workload $wl --steps 1 --warmup 3 --no-graph --skip-cpu --skip-roofline > $O/ncu_list_${wl}.log 2>&1
  echo "$wl rc=$?"; tail -1 $O/ncu_list_${wl}.log | cut -c1-200
done
