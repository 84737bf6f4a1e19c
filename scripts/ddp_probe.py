#!/usr/bin/env python
"""torchrun helper: cost of the gradient all-reduce alone (AVG vs SUM, eager vs graph) and per-rank step-time spread."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_900_000
buf = torch.randn(n, device=dev)
res = {}
for name, op in (("avg", dist.ReduceOp.AVG), ("sum", dist.ReduceOp.SUM)):
    for _ in range(5):
        dist.all_reduce(buf, op=op)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        dist.all_reduce(buf, op=op)
    e1.record(); torch.cuda.synchronize()
    res[name + "_eager_us"] = e0.elapsed_time(e1) * 1e3 / 50
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        dist.all_reduce(buf, op=op)
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(10):
            dist.all_reduce(buf, op=op)
    g.replay(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    res[name + "_graph_us"] = e0.elapsed_time(e1) * 1e3 / 50
    g.reset()
t = torch.tensor([res[k] for k in sorted(res)], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"world": world, "floats": n, **{k: round(v, 1) for k, v in zip(sorted(res), t.tolist())}}), flush=True)
dist.barrier()
dist.destroy_process_group()
