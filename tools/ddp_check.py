"""Multi-GPU row on real NCCL, run under torchrun:
  (1) correctness: each rank rolls out its contiguous shard of the start states (imagine fwd + hand-written BPTT),
      one flat NCCL all-reduce of the actor gradients -> must equal the single-process full-batch gradient;
  (2) timing of a BASELINE configs[4]-shaped step (discrete D=1024, 32x32, H=64, fwd + actor-BPTT bwd + gradient
      all-reduce), 8 192 start states per GPU by default (65 536 over 8 GPUs).
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ddp_check.py [--rows R] [--H H]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "world-model-rl_b200"), os.path.join(ROOT, "tests")]
import torch
import torch.distributed as dist
import helpers_gpu as HG
from wmrl.parallel import allreduce_gradients, reduce_scalar, shard_range

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=8192)
ap.add_argument("--H", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

def step(rssm, actor, init, noise, H, n_global):
    actor.zero_grad(set_to_none=True)
    seq = rssm.imagine_rollout(init, actor, H, noise=noise)
    feat = seq.get_features()
    loss = (feat ** 2).sum() / (n_global * H * feat.shape[-1])      # a mean over the GLOBAL batch
    loss.backward()
    n = allreduce_gradients(actor.parameters(), average=False)      # shards already carry 1/n_global
    return float(loss.detach()), n

# ---- (1) sharded == single process, small discrete model
variant, B, H, D, S, C, Hd, A = "discrete", 301, 6, 64, 8, 6, 48, 3
rssm, actor = HG.build_modules(variant, D, S, C, Hd, 16, A, 64, 2, seed=0)
for p in rssm.parameters(): p.requires_grad_(False)
d0, s0, nz = HG.synth_inputs(variant, B, H, D, S, C)
lo, hi = shard_range(B, rank, world)
init = HG.init_state(rssm, d0[lo:hi], s0[lo:hi])
loss_local, n_sent = step(rssm, actor, init, nz[:, lo:hi].contiguous().to(dev), H, B)
loss = reduce_scalar(loss_local, dev)
sharded = [None if p.grad is None else p.grad.detach().clone() for p in actor.parameters()]   # the sigma head gets no grad
full_init = HG.init_state(rssm, d0, s0)
actor.zero_grad(set_to_none=True)
seq = rssm.imagine_rollout(full_init, actor, H, noise=nz.to(dev))
feat = seq.get_features()
full_loss = (feat ** 2).sum() / (B * H * feat.shape[-1])
full_loss.backward()
worst = 0.0
for got, p in zip(sharded, actor.parameters()):
    if got is None or p.grad is None:
        assert got is None and p.grad is None
        continue
    scale = float(p.grad.abs().max()) + 1e-12
    worst = max(worst, float((got - p.grad).abs().max()) / scale)
ok = worst < 2e-4 and abs(loss - float(full_loss)) < 1e-5 * abs(float(full_loss)) + 1e-7
# ---- (2) configs[4] shape
variant, D, S, C, Hd, A = "discrete", 1024, 32, 32, 1024, 6
rssm, actor = HG.build_modules(variant, D, S, C, Hd, 16, A, 512, 3, seed=0)
for p in rssm.parameters(): p.requires_grad_(False)
R, Hh = args.rows, args.H
g = torch.Generator(device=dev).manual_seed(100 + rank)
d0 = torch.tanh(torch.randn(R, D, device=dev, generator=g))
s0 = torch.nn.functional.one_hot(torch.randint(0, S, (R, C), device=dev, generator=g), S).float().reshape(R, C * S)
nz = torch.empty(Hh, R, C, S, device=dev).exponential_(1.0, generator=g)
init = HG.init_state(rssm, d0, s0)
step(rssm, actor, init, nz, Hh, R * world)
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    _, n_sent = step(rssm, actor, init, nz, Hh, R * world)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
if rank == 0:
    fl = 45131776      # C5 fwd+bwd algorithmic FLOPs per latent step
    print(json.dumps({"check": "sharded actor gradient == single-process gradient (NCCL all-reduce)", "ok": bool(ok),
                      "max_rel_err": worst, "world": world, "grad_elements_allreduced": n_sent,
                      "c5_shape": {"variant": "discrete D1024 32x32 Hd1024 Actor 3x512", "rows_per_gpu": R, "H": Hh,
                                   "ms_per_step_fwd_bwd_allreduce": ms,
                                   "latent_steps_per_s": world * R * Hh / ms * 1e3,
                                   "tflops_per_gpu": fl * R * Hh / ms / 1e9,
                                   "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}}))
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
