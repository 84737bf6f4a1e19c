"""2 GPUs, NCCL: N ranks + GradSync (direct gradient sinks, one AVG all-reduce)  ==  1 rank on the concatenated
batch, for the FNO / TFNO CUDA path -- eager and as a whole-step CUDA graph (what bench.py replays at N > 1) --
and the process group tears down cleanly afterwards (no os._exit).

Skipped on single-GPU boxes; run with `gpurun --gpus 2 -- python -m pytest tests/test_ddp_nccl_gpu.py -m gpu`.
"""
import os
import socket
import threading

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
PER_RANK = 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(kind):
    import dlwp_benchmark_b200 as pkg
    torch.manual_seed(77)
    cls = pkg.TFNO if kind == "tfno" else pkg.FNO
    return cls(n_modes=(16, 16), hidden_channels=32, in_channels=1, out_channels=1, lifting_channels=256,
               projection_channels=256, n_layers=3, rank=0.8 if kind == "tfno" else 1.0)


def _data(world):
    g = torch.Generator().manual_seed(99)
    return torch.randn(PER_RANK * world, 1, 64, 64, generator=g), torch.randn(PER_RANK * world, 1, 64, 64, generator=g)


def _worker(rank, world, port, kind, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from dlwp_benchmark_b200.ddp import GradSync, shard_batch
    model = _model(kind).to(dev)
    params = list(model.parameters())
    sync = GradSync(params, world)                         # direct sinks on CUDA
    X, Y = _data(world)
    lo, hi = shard_batch(X.shape[0], rank, world)
    x, y = X[lo:hi].to(dev), Y[lo:hi].to(dev)

    def step():
        sync.zero()
        F.mse_loss(model(x), y).backward()
        sync.allreduce()

    step()
    step()                                                 # second step: zero() now skips the sink-written ranges
    torch.cuda.synchronize()
    eager = [p.grad.detach().clone().cpu() for p in params]
    used = [bool(getattr(p, "_sb200_sink_used", False)) for p in params]
    # whole step (zero + fwd + bwd + all-reduce) as ONE graph
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    sync.flat.fill_(7.0)                                   # poison: the replay must rewrite every gradient
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    graphed = [p.grad.detach().clone().cpu() for p in params]
    q.put((rank, eager, graphed, used))
    # clean teardown: drop the graph that captured the communicator first, then destroy the group
    watchdog = threading.Timer(60.0, lambda: os._exit(3))
    watchdog.daemon = True
    watchdog.start()
    dist.barrier()
    torch.cuda.synchronize()
    graph.reset()
    del graph
    torch.cuda.synchronize()
    dist.destroy_process_group()
    watchdog.cancel()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(300)
@pytest.mark.parametrize("kind", ["tfno", "fno"])
def test_two_ranks_equal_one_rank_on_the_full_batch(kind):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=90)
        assert p.exitcode == 0, f"rank exited with {p.exitcode} (3 = destroy_process_group hung)"
    # single rank, ordinary autograd accumulation, full batch
    dev = torch.device("cuda", 0)
    model = _model(kind).to(dev)
    X, Y = _data(world)
    F.mse_loss(model(X.to(dev)), Y.to(dev)).backward()
    ref = [p.grad.detach().cpu().double() for p in model.parameters()]
    names = [k for k, _ in model.named_parameters()]
    rel = lambda a, b: (torch.linalg.norm(a.double() - b) / torch.linalg.norm(b).clamp_min(1e-30)).item()
    worst = 0.0
    for r in range(world):
        for k, a, g, b in zip(names, res[r][1], res[r][2], ref):
            e1, e2 = rel(a, b), rel(g, b)
            worst = max(worst, e1, e2)
            assert e1 < 3e-6 and e2 < 3e-6, (k, e1, e2)
    for a, b in zip(res[0][1], res[1][1]):
        assert torch.equal(a, b)                           # ranks hold bit-identical reduced gradients
    for a, b in zip(res[0][1], res[0][2]):
        assert torch.equal(a, b)                           # graph replay == eager
    assert all(res[0][3]), [k for k, u in zip(names, res[0][3]) if not u]   # every gradient was written in place
    print(f"{kind}: worst rel-L2 of a reduced gradient vs the 1-rank full-batch gradient = {worst:.2e}")
