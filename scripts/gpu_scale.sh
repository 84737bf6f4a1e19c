#!/bin/bash
# Weak-scaling check on one box: N = 1, 2, 8 back to back (driver does 1, 2, 4, 8 at round end).
O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -8
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --skip-cpu --skip-roofline > $O/scale_n1.json 2> $O/scale_n1.err; cut -c1-260 $O/scale_n1.json
for n in 2 8; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 30 --warmup 5 --skip-cpu --skip-roofline > $O/scale_n$n.json 2> $O/scale_n$n.err
  echo "n=$n rc=$?"; grep '^{"metric"' $O/scale_n$n.json | cut -c1-260; grep -i "capture failed\|error\|NVLS" $O/scale_n$n.err | head -5
done
python - <<PY
import json
v={}
for n in (1,2,8):
    try:
        for l in open('$O/scale_n%d.json'%n):
            if l.startswith('{"metric"'): d=json.loads(l); v[n]=(d['value'],d['ms_per_step'],d['config'].get('graph_scope'))
    except Exception as e: print(n,'ERR',e)
print(v)
if 1 in v:
    for n in v: print(n, 'speedup %.2f'%(v[n][0]/v[1][0]), 'efficiency %.3f'%(v[n][0]/v[1][0]/n))
PY
