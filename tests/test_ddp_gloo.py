"""CPU, world_size 2, gloo: the data-parallel harness (GradSync) -- the N>1 path of bench.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dlwp_benchmark_b200.ddp import GradSync, shard_batch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                   # identical weights on every rank
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    unused = torch.nn.Parameter(torch.ones(4))             # e.g. AFNONet.norm: never receives a gradient
    params = list(model.parameters()) + [unused]
    sync = GradSync(params, world, bucket_bytes=64)        # tiny buckets -> several all-reduces
    g = torch.Generator().manual_seed(123)
    X, Y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    lo, hi = shard_batch(8, rank, world)
    sync.zero()
    torch.nn.functional.mse_loss(model(X[lo:hi]), Y[lo:hi]).backward()
    sync.allreduce()
    q.put((rank, [p.grad.clone() for p in params], len(sync.buckets)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradsync_matches_full_batch_gradient():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=90) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    # single-process reference on the full batch (equal shard sizes => mean of shard grads == full-batch grad)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    g = torch.Generator().manual_seed(123)
    X, Y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    torch.nn.functional.mse_loss(model(X), Y).backward()
    ref = [p.grad for p in model.parameters()] + [torch.zeros(4)]
    assert res[0][2] > 1
    for r in range(world):
        for a, b in zip(res[r][1], ref):
            assert torch.allclose(a, b, atol=1e-6, rtol=1e-5)
    for a, b in zip(res[0][1], res[1][1]):
        assert torch.equal(a, b)                           # ranks hold bit-identical reduced gradients


def test_shard_batch_covers_everything_once():
    for n, w in ((512, 8), (4096, 8), (10, 4), (3, 8)):
        spans = [shard_batch(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0]
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
