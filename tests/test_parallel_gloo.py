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


# --------------------------------------------------------------------------------------------------
# the RSSM path itself, sharded: each rank rolls out its contiguous slice of the start states (here through
# the CPU oracle - the product path has no CPU implementation), back-propagates the actor loss and joins the
# single flat gradient all-reduce.  The result must equal the full-batch gradient: the data-parallel layout of
# SURVEY 8e (rows independent, weights replicated, one collective) written down as a test.
# --------------------------------------------------------------------------------------------------
def _rssm_problem():
    import wmrl
    torch.manual_seed(0)
    D, S, A, Ha, La, B, H = 24, 6, 2, 32, 2, 13, 4
    rssm = wmrl.ContinuousRSSM(D, D, S, 16, A)
    actor = wmrl.Actor(D + S, A, 1.0, La, Ha)
    g = torch.Generator().manual_seed(3)
    deter0, stoch0 = torch.tanh(torch.randn(B, D, generator=g)), torch.randn(B, S, generator=g)
    noise = torch.randn(H, B, S, generator=g)
    gfeat = torch.randn(B, H, D + S, generator=g) / B
    return rssm, actor, deter0, stoch0, noise, gfeat, (S, H)


def _rssm_actor_grads(rssm, actor, deter0, stoch0, noise, gfeat, S, H):
    from oracle import rssm_oracle as O
    w = {k: v.detach() for k, v in rssm.state_dict().items()}
    aw = dict(actor.named_parameters())
    out = O.imagine_rollout("continuous", w, aw, deter0, stoch0, noise, H, S)
    (torch.cat([out["deter"], out["stoch"]], -1) * gfeat).sum().backward()


def _rssm_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rssm, actor, deter0, stoch0, noise, gfeat, (S, H) = _rssm_problem()
        lo, hi = shard_range(deter0.shape[0], rank, world)
        _rssm_actor_grads(rssm, actor, deter0[lo:hi], stoch0[lo:hi], noise[:, lo:hi], gfeat[lo:hi], S, H)
        n = allreduce_gradients([p for p in actor.parameters() if p.grad is not None], average=False)
        if rank == 0:
            out.put((n, {k: p.grad.clone() for k, p in actor.named_parameters() if p.grad is not None}))
    finally:
        dist.destroy_process_group()


def test_sharded_rssm_actor_gradients_match_full_batch():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rssm_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    n, grads = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rssm, actor, deter0, stoch0, noise, gfeat, (S, H) = _rssm_problem()
    _rssm_actor_grads(rssm, actor, deter0, stoch0, noise, gfeat, S, H)
    ref = {k: p.grad for k, p in actor.named_parameters() if p.grad is not None}
    assert set(ref) == set(grads) and n == sum(v.numel() for v in ref.values())
    for k in ref:
        torch.testing.assert_close(grads[k], ref[k], rtol=1e-4, atol=1e-7)
