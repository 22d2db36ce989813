import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "world-model-rl_b200")]
import torch, wmrl
sys.argv = [sys.argv[0]]
import bench
c = dict(bench.CFG); dev = torch.device("cuda")
rssm, actor = bench.build_problem(c, dev)
d0, s0, nz = [t.to(dev) for t in bench.host_inputs(c, c["B"], 1, False)]
init = wmrl.InternalStateContinuous(d0, s0, s0.clone(), torch.ones_like(s0))
def step():
    with torch.no_grad():
        return rssm.imagine_rollout(init, actor, c["H"], noise=nz)
for _ in range(10): seq = step()
torch.cuda.synchronize()
st0 = torch.cuda.memory_stats()
host, gpu = [], []
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(60)]
for i in range(60):
    t0 = time.perf_counter(); ev[i][0].record(); seq = step(); ev[i][1].record(); host.append((time.perf_counter() - t0) * 1e3)
torch.cuda.synchronize()
st1 = torch.cuda.memory_stats()
print("host enqueue ms:", " ".join(f"{h:.2f}" for h in host))
print("event ms      :", " ".join(f"{a.elapsed_time(b):.1f}" for a, b in ev))
for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "allocation.all.allocated"):
    print(k, st1.get(k, 0) - st0.get(k, 0))
print("load avg", os.getloadavg(), "cpus", os.cpu_count())
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(50): seq = step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
