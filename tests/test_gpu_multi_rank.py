"""Two ranks on two CUDA devices (NCCL): the sharded wmrl training step - bf16 rollout + persistent BPTT per
rank, one flat all-reduce of the actor gradients - reproduces the single-process full-batch gradient.
Skipped on boxes with fewer than two GPUs (the driver's 1-GPU test tier); tools/ddp_check.py runs the same
check at 2/4/8 ranks."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(dev, precision):
    import helpers_gpu as HG
    D, S, Hd, A, Ha, La, B, H = 200, 30, 200, 6, 512, 3, 700, 6
    rssm, actor = HG.build_modules("continuous", D, S, 1, Hd, 64, A, Ha, La, seed=0, device=dev, precision=precision)
    for p in rssm.parameters():
        p.requires_grad_(False)
    d0, s0, nz = HG.synth_inputs("continuous", B, H, D, S, 1)
    gf = torch.randn(B, H, D + S, generator=torch.Generator().manual_seed(5)) / B
    return HG, rssm, actor, d0, s0, nz, gf, H


def _grads(HG, rssm, actor, d0, s0, nz, gf, H, dev):
    init = HG.init_state(rssm, d0, s0, device=dev)
    seq = rssm.imagine_rollout(init, actor, H, noise=nz.to(dev))
    (seq.get_features() * gf.to(dev)).sum().backward()


def _worker(rank, world, port, precision, out):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    for p in (root, os.path.join(root, "world-model-rl_b200"), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from wmrl.parallel import allreduce_gradients, shard_range
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = torch.device("cuda", rank)
    # deliberately NOT torch.cuda.set_device(rank): the native calls must follow the tensors' device
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        HG, rssm, actor, d0, s0, nz, gf, H = _problem(dev, precision)
        lo, hi = shard_range(d0.shape[0], rank, world)
        _grads(HG, rssm, actor, d0[lo:hi], s0[lo:hi], nz[:, lo:hi].contiguous(), gf[lo:hi], H, dev)
        allreduce_gradients([p for p in actor.parameters() if p.grad is not None], average=False)
        if rank == 0:
            out.put({k: p.grad.cpu() for k, p in actor.named_parameters() if p.grad is not None})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_two_rank_sharded_training_step_matches_single_process(precision):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, precision, out)) for r in range(2)]
    for p in procs:
        p.start()
    grads = out.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    HG, rssm, actor, d0, s0, nz, gf, H = _problem("cuda:0", precision)
    _grads(HG, rssm, actor, d0, s0, nz, gf, H, "cuda:0")
    for k, p in actor.named_parameters():
        if p.grad is None:
            continue
        ref = p.grad.cpu()
        rel = float((grads[k] - ref).norm() / (ref.norm() + 1e-30))
        assert rel < (2e-3 if precision == "bf16" else 1e-4), f"{k}: {rel:.3e}"


def test_native_calls_follow_the_tensors_device():
    """modules on cuda:1 while cuda:0 is torch's current device (the reference handles this): the
    constants, function attributes and launches of libwmrl must land on cuda:1."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import helpers_gpu as HG
    from oracle import rssm_oracle as O
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 1)
    D, S, A = 200, 30, 6
    rssm, actor = HG.build_modules("continuous", D, S, 1, D, 64, A, 512, 3, seed=1, device=dev, precision="bf16")
    d0, s0, nz = HG.synth_inputs("continuous", 200, 4, D, S, 1)
    with torch.no_grad():
        seq = rssm.imagine_rollout(HG.init_state(rssm, d0, s0, device=dev), actor, 4, noise=nz.to(dev))
        ref = O.imagine_rollout("continuous", HG.cpu_weights(rssm), HG.cpu_weights(actor), d0, s0, nz, 4, S)
    assert seq.deter_states.device == dev
    assert float((seq.deter_states[:, 1:].cpu() - ref["deter"]).abs().max()) < 2e-2
