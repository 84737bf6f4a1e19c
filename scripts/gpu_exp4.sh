#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest_gpu_s5c.log; cat $O/pytest_gpu_s5c.log
echo "== kbench cfg3 (tc analysis / tc row only / ffma)"
timeout 300 python scripts/kbench.py cfg3 | cut -c1-300
SB200_TC_COLDFT_OFF=1 timeout 300 python scripts/kbench.py cfg3 | cut -c1-300
echo "== bench cfg3 / cfg5"
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --skip-cpu > $O/bench_cfg3_s5c.json 2>/dev/null; python -c "import sys,json; d=json.load(open('$O/bench_cfg3_s5c.json')); print('cfg3 ms/step', d['ms_per_step'], d['roofline']['all'])"
timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 3 > $O/bench_cfg5_s5c.json 2>/dev/null; python -c "import sys,json; d=json.load(open('$O/bench_cfg5_s5c.json')); print('cfg5 ms/model step', d['config']['ms_per_model_step'], d['roofline']['frac'])"
