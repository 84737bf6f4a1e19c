#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "rowdft or analysis or stages or linearity" 2>&1 | tail -15 > $O/pytest_rowdft.log; cat $O/pytest_rowdft.log
echo "== kbench cfg3 (tc rowdft on / off)"
timeout 300 python scripts/kbench.py cfg3 | cut -c1-600
SB200_TC_ROWDFT_OFF=1 timeout 300 python scripts/kbench.py cfg3 | cut -c1-600
echo "== bench cfg3"
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --skip-cpu --skip-roofline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg3 ms/step', d['ms_per_step'])"
SB200_TC_ROWDFT_OFF=1 timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --skip-cpu --skip-roofline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg3 (ffma rowdft) ms/step', d['ms_per_step'])"
