#!/bin/bash
# Run on the GPU box (via gpurun): short bench (eager + graph) and the ncu launch list.
mkdir -p gpurun_out
WL=${1:-cfg2}
timeout 600 python bench.py --workload $WL --steps 10 --warmup 3 --no-graph --skip-cpu > gpurun_out/bench_${WL}_eager.json 2> gpurun_out/bench_${WL}_eager.err
tail -c 3000 gpurun_out/bench_${WL}_eager.json; tail -5 gpurun_out/bench_${WL}_eager.err
timeout 600 python bench.py --workload $WL --steps 20 --warmup 5 > gpurun_out/bench_${WL}.json 2> gpurun_out/bench_${WL}.err
tail -c 3000 gpurun_out/bench_${WL}.json; tail -5 gpurun_out/bench_${WL}.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${WL}.csv \
   python bench.py --workload $WL --steps 1 --warmup 3 --no-graph --skip-cpu > gpurun_out/ncu_${WL}.log 2>&1
tail -3 gpurun_out/ncu_${WL}.log
