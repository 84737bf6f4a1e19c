#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "rowdft or analysis or stages or linearity or tc_fno_block" 2>&1 | tail -6
timeout 300 python scripts/kbench.py cfg3 | cut -c1-200
timeout 300 python scripts/kbench.py cfg5x 2>/dev/null | cut -c1-200
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --skip-cpu --skip-roofline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg3 ms/step', d['ms_per_step'])"
