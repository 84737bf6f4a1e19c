#!/bin/bash
# 8-GPU job: BASELINE configs[2] (FNO2D 256x256, global batch 512 = 64/GPU x 8, data parallel) and
# configs[4] (100-step rollout, 4096 initial conditions = 512/GPU x 8, sharded, no collective), plus cfg2 at N=8.
O=gpurun_out; mkdir -p $O
run() { # workload N port steps
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 bench.py \
     --workload $1 --gpus $2 --steps $4 --warmup 3 --skip-cpu > $O/r02_bench_$1_n$2.json 2> $O/r02_bench_$1_n$2.err
  echo "$1 N=$2 rc=$?"; cut -c1-260 $O/r02_bench_$1_n$2.json
}
run cfg3 8 29521 20
run cfg5 8 29522 2
run cfg2 8 29523 100
timeout 200 python bench.py --workload cfg3 --gpus 1 --steps 20 --warmup 3 --skip-cpu --skip-roofline > $O/r02_bench_cfg3_n1.json 2> $O/r02_bench_cfg3_n1.err; cut -c1-260 $O/r02_bench_cfg3_n1.json
timeout 200 python bench.py --workload cfg5 --gpus 1 --steps 2 --warmup 3 --skip-cpu > $O/r02_bench_cfg5_n1.json 2> $O/r02_bench_cfg5_n1.err; cut -c1-260 $O/r02_bench_cfg5_n1.json
grep -l "did not return" $O/r02_bench_*_n8.err
