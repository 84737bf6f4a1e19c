#!/bin/bash
# 2-GPU check of the final round-2 code: NCCL equivalence tests + cfg2 / cfg4 bench at N=2 (every multi-rank command under timeout)
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_ddp_nccl_gpu.py tests/test_parity_configs_gpu.py -m gpu -q -k "two_ranks or non_current" 2>&1 | tail -5 > $O/r02f_pytest_2gpu.log
tail -3 $O/r02f_pytest_2gpu.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --skip-roofline > $O/r02f_bench_cfg2_n2.json 2> $O/r02f_bench_cfg2_n2.err; echo "n2 rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg4 --steps 10 --warmup 3 --skip-roofline > $O/r02f_bench_cfg4_n2.json 2> $O/r02f_bench_cfg4_n2.err; echo "cfg4 n2 rc=$?"
cut -c1-260 $O/r02f_bench_cfg2_n2.json; cut -c1-200 $O/r02f_bench_cfg4_n2.json; tail -2 $O/r02f_bench_cfg4_n2.err
