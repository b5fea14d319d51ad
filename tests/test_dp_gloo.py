"""World-size-2 gloo test (CPU) of the data-parallel host logic: one flat gradient bucket, all-reduce mean."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sod100k_b200.trainer import FlatGrads

    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 1, 3, 3))]
    flat = FlatGrads(params)
    # autograd accumulates IN PLACE into the bucket views
    loss = sum(((rank + 1.0) * (i + 1) * p).sum() for i, p in enumerate(params))
    loss.backward()
    assert flat.intact()
    flat.all_reduce_mean()
    want = sum(r + 1.0 for r in range(world)) / world
    ok = all(torch.allclose(p.grad, torch.full_like(p, want * (i + 1))) for i, p in enumerate(params))
    ok = ok and flat.bucket.numel() == 12 + 7 + 18 and flat.intact()
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
