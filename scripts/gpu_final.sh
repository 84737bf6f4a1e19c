#!/bin/bash
# Evidence pass: parity tests, smoke, every bench workload, the reference arm, ncu launch lists and --set full captures.
TAG=${1:-s5f}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu_$TAG.txt 2>&1; (nproc; free -g | head -2) >> $O/gpu_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest_gpu_$TAG.log; tail -3 $O/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; tail -1 $O/smoke_$TAG.log
timeout 600 python bench.py > $O/bench_cfg2_$TAG.json 2> $O/bench_cfg2_$TAG.err; echo "bench rc=$?"; cut -c1-400 $O/bench_cfg2_$TAG.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref_cfg2_$TAG.json 2>/dev/null; cut -c1-200 $O/bench_ref_cfg2_$TAG.json
timeout 300 python bench.py --tc-mode 1 --skip-cpu > $O/bench_cfg2_tf32_$TAG.json 2>/dev/null
for wl in cfg1 cfg2d cfg3 cfg4; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --skip-cpu > $O/bench_${wl}_$TAG.json 2>/dev/null
done
timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 3 > $O/bench_cfg5_$TAG.json 2>/dev/null
python - <<PY
import json,glob
for f in sorted(glob.glob('$O/bench_*_$TAG.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], d.get('ms_per_step'), d.get('value'), (d.get('roofline') or {}).get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
for wl in cfg2 cfg3; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_${wl}_$TAG.csv \
   python bench.py --workload $wl --steps 1 --warmup 3 --no-graph --skip-cpu --skip-roofline > /dev/null 2>&1
echo "ncu list $wl rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:tc_pointwise|tc_wgrad_kernel|analysis_fused|modes_gemm|cgemm_kernel|cskinny|coldft' -c 48 -f -o $O/full_cfg2_$TAG \
   python bench.py --workload cfg2 --steps 1 --warmup 3 --no-graph --skip-cpu --skip-roofline > /dev/null 2>&1
ncu -i $O/full_cfg2_$TAG.ncu-rep --page raw --csv > $O/full_cfg2_${TAG}_raw.csv 2>/dev/null; rm -f $O/full_cfg2_$TAG.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:tc_rowdft|coldft_inv|tc_pointwise_kernel<\(int\)3, \(int\)[13]|tc_wgrad_kernel' -c 14 -f -o $O/full_cfg3_$TAG \
   python bench.py --workload cfg3 --steps 1 --warmup 3 --no-graph --skip-cpu --skip-roofline > /dev/null 2>&1
ncu -i $O/full_cfg3_$TAG.ncu-rep --page raw --csv > $O/full_cfg3_${TAG}_raw.csv 2>/dev/null
ncu -i $O/full_cfg3_$TAG.ncu-rep --page source --csv -k regex:tc_rowdft -c 2 > $O/full_cfg3_${TAG}_src_rowdft.csv 2>/dev/null
rm -f $O/full_cfg3_$TAG.ncu-rep
du -sh $O
