#!/bin/bash
mkdir -p gpurun_out
(nproc; free -g; ulimit -a) > gpurun_out/host.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | tail -15
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 5 --skip-cpu > gpurun_out/bench_cfg2_graph.json 2> gpurun_out/bench_cfg2_graph.err
echo "rc=$?"; tail -c 1500 gpurun_out/bench_cfg2_graph.json; tail -8 gpurun_out/bench_cfg2_graph.err
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-graph > gpurun_out/bench_cfg2_cpu.json 2> gpurun_out/bench_cfg2_cpu.err
echo "rc=$?"; tail -c 600 gpurun_out/bench_cfg2_cpu.json; tail -8 gpurun_out/bench_cfg2_cpu.err
