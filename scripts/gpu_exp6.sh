#!/bin/bash
O=gpurun_out; mkdir -p $O
SB200_ANALYSIS_PREFER=tc timeout 600 python -m pytest tests -m gpu -x -q -k "analysis or cfg2 or golden or cfg1" 2>&1 | tail -4
timeout 300 python scripts/kbench.py cfg2 | cut -c1-160
SB200_ANALYSIS_PREFER=tc timeout 300 python scripts/kbench.py cfg2 | cut -c1-160
for v in "X=1" "SB200_ANALYSIS_PREFER=tc"; do
env $v timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-roofline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v cfg2 ms/step', d['ms_per_step'])"
done
