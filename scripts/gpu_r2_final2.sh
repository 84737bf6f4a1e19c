#!/bin/bash
# Round-2 final evidence (subset of gpu_r2_final.sh, ~6 GPU-minutes): parity suite, smoke, the driver-format bench line and
# reference arm, the other workloads without CPU baselines, launch list + `--set full` table of the cfg2 step.
TAG=${1:-r02f}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/${TAG}_pytest_gpu.log; tail -2 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 600 python bench.py --steps 200 --warmup 5 > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_ref_cfg2.json 2>/dev/null
timeout 300 python bench.py --tc-mode 1 --steps 100 --skip-cpu > $O/${TAG}_bench_cfg2_tf32.json 2>/dev/null
for wl in cfg1 cfg2d cfg3 cfg4 cfg4f; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --skip-cpu > $O/${TAG}_bench_${wl}.json 2>/dev/null
done
timeout 600 python bench.py --workload cfg4 --steps 20 --warmup 3 --tc-mode 1 --skip-cpu > $O/${TAG}_bench_cfg4_tf32.json 2>/dev/null
timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 3 --skip-cpu > $O/${TAG}_bench_cfg5.json 2>/dev/null
python - <<PY
import json,glob
for f in sorted(glob.glob('$O/${TAG}_bench_*.json')):
    try:
        d=json.load(open(f)); r=d.get('roofline') or {}
        print(f.split('/')[-1], 'ms', round(d.get('ms_per_step'),3), 'value', round(d.get('value'),1), 'roof', r.get('kernel'), r.get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
for wl in cfg2 cfg4; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches_${wl}.csv \
   python bench.py --workload $wl --steps 1 --warmup 3 --no-graph --skip-cpu --skip-roofline > /dev/null 2>&1
python scripts/summarize_launches.py $O/${TAG}_launches_${wl}.csv > $O/${TAG}_launches_${wl}.txt 2>&1; rm -f $O/${TAG}_launches_${wl}.csv
done
timeout 1200 ncu --set full --clock-control none -k 'regex:tc_pointwise|tc_wgrad_kernel|analysis_fused|modes_gemm|cgemm_kernel|cskinny|coldft|tc_rowdft' -c 70 -f -o /tmp/full_cfg2 \
   python bench.py --workload cfg2 --steps 1 --warmup 3 --no-graph --skip-cpu --skip-roofline > /dev/null 2>&1
ncu -i /tmp/full_cfg2.ncu-rep --page raw --csv > /tmp/full_cfg2_raw.csv 2>/dev/null
python scripts/ncu_raw_summary.py /tmp/full_cfg2_raw.csv $O/${TAG}_cfg2_ncu_full "ncu --set full --clock-control none, first eager train step of bench.py --workload cfg2 (round 2, final code)" > /dev/null
timeout 900 ncu --set full --clock-control none -k 'regex:tc_gemm|ln_|blocklinear|cl_|colsum|rows_reduce|afno_|mask_mul' -c 80 -f -o /tmp/full_cfg4 \
   python bench.py --workload cfg4 --steps 1 --warmup 3 --no-graph --skip-cpu > /dev/null 2>&1
ncu -i /tmp/full_cfg4.ncu-rep --page raw --csv > /tmp/full_cfg4_raw.csv 2>/dev/null
python scripts/ncu_raw_summary.py /tmp/full_cfg4_raw.csv $O/${TAG}_cfg4_ncu_full "ncu --set full --clock-control none, first eager train step of bench.py --workload cfg4 (round 2, final code)" > /dev/null
du -sh $O
