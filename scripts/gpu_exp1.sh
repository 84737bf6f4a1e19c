#!/bin/bash
# Experiment pass 1: PDL on/off, ring depth, raw-hi operand, streaming microbenchmarks.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/pytest_pdl1.log; cat $O/pytest_pdl1.log
SB200_TC_RAWHI=1 timeout 900 python -m pytest tests -m gpu -q -k "tc_ or cfg2 or fused or lift or golden or cfg1" 2>&1 | tail -12 > $O/pytest_rawhi.log; cat $O/pytest_rawhi.log
i=0
for v in "SB200_PDL=0 SB200_TP_SMEM_KB=208" "SB200_PDL=1 SB200_TP_SMEM_KB=208" "SB200_PDL=1" "SB200_PDL=0" "SB200_PDL=1 SB200_TC_RAWHI=1" "SB200_PDL=1 SB200_TC_MODE=1"; do
  i=$((i+1))
  echo "== $v"
  env $v timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-roofline 2> $O/exp1_bench_$i.err | tee $O/exp1_bench_$i.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['config'].get('graph_scope'))"
  tail -2 $O/exp1_bench_$i.err
done
echo "== kbench"
timeout 300 python scripts/kbench.py cfg2 | tee $O/exp1_kbench_default.json
SB200_TC_RAWHI=1 timeout 300 python scripts/kbench.py cfg2 | tee $O/exp1_kbench_rawhi.json
SB200_TP_SMEM_KB=208 timeout 300 python scripts/kbench.py cfg2 | tee $O/exp1_kbench_s2.json
echo "== membench"
timeout 300 bash tools/membench.sh 2>&1 | tee $O/exp1_membench.txt
echo "== wgrad"
for v in "X=1" "SB200_WG_DEBUG=1" "SB200_WG_CH=2" "SB200_WG_CH=2 SB200_WG_STAGES=3" "SB200_WG_STAGES=3" "SB200_TC_RAWHI=1" "SB200_TC_MODE=1"; do
  env $v timeout 120 python scripts/kb_wgrad.py 2>&1 | tail -1 | tee -a $O/exp1_wgrad.txt
done
