"""N>1 host logic on CPU: world_size 2, gloo.  Batch sharding + the single flat gradient
all-reduce reproduce the single-process gradient of a mean loss."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wmrl.parallel import allreduce_gradients, reduce_scalar, shard_range, shard_rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                      # replicated weights
        model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
        g = torch.Generator().manual_seed(1)
        x = torch.randn(n_rows, 6, generator=g)   # the global batch, sharded by rows
        xs = shard_rows(x, rank, world)
        lo, hi = shard_range(n_rows, rank, world)
        assert xs.shape[0] == hi - lo
        loss = model(xs).mean()                   # local mean over the shard
        loss.backward()
        n = allreduce_gradients(model.parameters(), weight=(hi - lo) / n_rows)
        total_rows = reduce_scalar(float(hi - lo), "cpu")
        if rank == 0:
            out.put((n, total_rows, [p.grad.clone() for p in model.parameters()]))
    finally:
        dist.destroy_process_group()


def test_sharded_gradients_match_single_process():
    world, n_rows = 2, 11                         # uneven shards: 6 + 5 rows
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rows, out)) for r in range(world)]
    for p in procs:
        p.start()
    n, total_rows, grads = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    x = torch.randn(n_rows, 6, generator=torch.Generator().manual_seed(1))
    model(x).mean().backward()
    assert total_rows == n_rows
    assert n == sum(p.numel() for p in model.parameters())
    for got, p in zip(grads, model.parameters()):
        torch.testing.assert_close(got, p.grad, rtol=1e-5, atol=1e-7)
