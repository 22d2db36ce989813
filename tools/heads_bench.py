"""Reward / done models + critic + target critic over the imagined features, fp32 (3xTF32 GEMM + fused LayerNorm
kernels) vs the bf16 tcgen05 chain: forward and the value-trainer's forward+backward.  usage: python tools/heads_bench.py [B H]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "world-model-rl_b200"))
import torch
import wmrl

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
H = int(sys.argv[2]) if len(sys.argv) > 2 else 15
F_ = 230
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(B, H, F_, device=dev)
out = {"B": B, "H": H}
MAC = {"critic": F_ * 512 + 2 * 512 * 512 + 512, "dense": F_ * 256 + 2 * 256 * 256 + 256}
fwd_flop = 2 * (2 * MAC["critic"] + 2 * MAC["dense"])
# backward: critic dgrad + wgrad, target critic dgrad, reward dgrad (done is thresholded: no gradient)
bwd_flop = 2 * (3 * MAC["critic"] + MAC["dense"])
for prec in ("bf16", "fp32"):
    reward = wmrl.DenseModel(input_dim=F_, precision=prec).to(dev)
    done = wmrl.DenseModel(input_dim=F_, output_act=torch.nn.Sigmoid, precision=prec).to(dev)
    critic = wmrl.ValueCritic(F_, precision=prec).to(dev)
    target = wmrl.ValueCritic(F_, precision=prec).to(dev)
    for m in (reward, done, target):
        for p in m.parameters():
            p.requires_grad_(False)

    def fwd():
        with torch.no_grad():
            return reward(x), done(x), critic(x), target(x)

    def train():
        xs = x.detach().requires_grad_(True)
        r, d, v, tv = reward(xs), done(xs), critic(xs), target(xs)
        d = wmrl.functional.done_mask(d)
        tgt = wmrl.functional.lambda_return(tv, r, d, 0.99, 0.95)
        loss = (0.5 * (tgt.detach() - v) ** 2).mean() - tgt.mean()
        loss.backward()
        return loss

    for name, fn, flop in (("fwd", fwd, fwd_flop), ("fwd_bwd", train, fwd_flop + bwd_flop)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out[f"{prec}_{name}_ms"] = ms
        out[f"{prec}_{name}_tflops"] = flop * B * H / ms / 1e9
    out[f"{prec}_peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2**30
    del reward, done, critic, target
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
print(json.dumps(out))
