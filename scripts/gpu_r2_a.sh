#!/bin/bash
# Round 2, GPU job A: probe, parity suite, cuFFT reproducer, bench (cfg2), sanitizer smoke.
O=gpurun_out; mkdir -p $O
bash scripts/probe_neuralop.sh > $O/r02_neuralop_probe.txt 2>&1
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > $O/r02_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > $O/r02_pytest_gpu_a.log
python scripts/cufft_vs_cpu_irfftn.py > $O/r02_cufft_vs_cpu.jsonl 2> $O/r02_cufft_vs_cpu.err
timeout 600 python bench.py --steps 200 --warmup 5 > $O/r02_bench_cfg2_a.json 2> $O/r02_bench_cfg2_a.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_ref_cfg2_a.json 2> $O/r02_bench_ref_cfg2_a.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitizer_smoke.py > $O/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitizer_smoke.py > $O/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/r02_sanitizer_racecheck.log
tail -5 $O/r02_pytest_gpu_a.log; cat $O/r02_bench_cfg2_a.json | cut -c1-600; tail -3 $O/r02_sanitizer_memcheck.log $O/r02_sanitizer_racecheck.log
