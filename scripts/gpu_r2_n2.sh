#!/bin/bash
# 2-GPU job: NCCL equivalence test (2 ranks == 1 rank on the full batch, eager + graph, clean teardown), non-current
# device test, cfg2 bench at N=2 with a clean destroy_process_group.  Every multi-rank command under `timeout`.
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_ddp_nccl_gpu.py tests/test_parity_configs_gpu.py -m gpu -q -s -k "two_ranks or non_current" 2>&1 | tail -25 > $O/r02_pytest_2gpu.log
tail -6 $O/r02_pytest_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > $O/r02_bench_cfg2_n2.json 2> $O/r02_bench_cfg2_n2.err; echo "n2 rc=$?"
timeout 200 python bench.py --gpus 1 --steps 100 --warmup 5 --skip-cpu --skip-roofline > $O/r02_bench_cfg2_n1.json 2> $O/r02_bench_cfg2_n1.err
cut -c1-230 $O/r02_bench_cfg2_n2.json; cut -c1-230 $O/r02_bench_cfg2_n1.json; tail -3 $O/r02_bench_cfg2_n2.err
python scripts/cufft_vs_cpu_irfftn.py > $O/r02_cufft_vs_cpu.jsonl 2> $O/r02_cufft_vs_cpu.err
