#!/bin/bash
# Pass 2: parity tests after the PDL removal, analysis ring A/B, all workloads of bench.py.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest_gpu_s5b.log; cat $O/pytest_gpu_s5b.log
echo "== kbench analysis ring"
for v in "SB200_AF_RING=22" "SB200_AF_RING=31" "SB200_AF_RING=21" "X=0"; do
  env $v timeout 200 python scripts/kbench.py cfg2 | cut -c1-400
done
echo "== bench cfg2 (full line)"
timeout 600 python bench.py --steps 30 --warmup 5 > $O/bench_cfg2_s5b.json 2> $O/bench_cfg2_s5b.err; tail -c 2500 $O/bench_cfg2_s5b.json; tail -2 $O/bench_cfg2_s5b.err
SB200_AF_RING=22 timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-roofline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('AF_RING=22', d['ms_per_step'])"
timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-roofline --tc-mode 1 2>/dev/null > $O/bench_cfg2_tf32_s5b.json; python -c "import sys,json; d=json.load(open('$O/bench_cfg2_tf32_s5b.json')); print('tc-mode 1', d['ms_per_step'])"
for wl in cfg1 cfg2d cfg3 cfg4; do
  echo "== bench $wl"
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --skip-cpu > $O/bench_${wl}_s5b.json 2> $O/bench_${wl}_s5b.err; tail -c 1800 $O/bench_${wl}_s5b.json; tail -2 $O/bench_${wl}_s5b.err
done
echo "== bench cfg5"
timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 3 > $O/bench_cfg5_s5b.json 2> $O/bench_cfg5_s5b.err; tail -c 1800 $O/bench_cfg5_s5b.json; tail -3 $O/bench_cfg5_s5b.err
