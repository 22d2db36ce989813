"""Timeline of the tcgen05 kernel's jobs (CTA 0, first tile, first steps): per job the SM-clock
offsets of MMA start / MMA issued+committed / epilogue wake / epilogue done."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "world-model-rl_b200")]
import torch
sys.argv = [sys.argv[0]]
import bench
from wmrl import _lib, functional as Fn

c = dict(bench.CFG); c["B"] = int(os.environ.get("B", 65536))
dev = torch.device("cuda")
rssm, actor = bench.build_problem(c, dev)
d0, s0, nz = [t.to(dev) for t in bench.host_inputs(c, c["B"], 1, False)]
lib = _lib.load()
La, Ha, actor_p = __import__("wmrl.rssm", fromlist=["x"])._actor_tensors(actor)
cfg = rssm._config(c["A"], Ha, La, 5.0)
bound = torch.ones(c["A"], device=dev)
rssm_p = rssm._rssm_tensors()
packed = rssm._packed.get(cfg, rssm_p, actor_p, bound)
B, H = c["B"], c["H"]
deter = torch.empty(B, H + 1, c["D"], device=dev); stoch = torch.empty(B, H + 1, c["S"], device=dev)
mean = torch.empty_like(stoch); std = torch.empty_like(stoch); act = torch.empty(B, H, c["A"], device=dev)
d = cfg.dims(B, H, 4, _lib.BF16)
ws_n = lib.wmrl_workspace_bytes(C.byref(d), 0)
ws = torch.zeros(ws_n, dtype=torch.uint8, device=dev)
w, aw = Fn._rssm_ptrs([p.detach() for p in rssm_p]), Fn._actor_ptrs([p.detach() for p in actor_p], La, bound)
for _ in range(2):
    ws.zero_()
    _lib.check(lib.wmrl_imagine_fwd(C.byref(d), C.byref(w), C.byref(aw), packed.data_ptr(), d0.data_ptr(), s0.data_ptr(),
                                    nz.data_ptr(), deter.data_ptr(), stoch.data_ptr(), mean.data_ptr(), std.data_ptr(),
                                    act.data_ptr(), None, None, 0, ws.data_ptr(), ws_n, torch.cuda.current_stream().cuda_stream), "fwd")
torch.cuda.synchronize()
njobs = 3 + 1 + 1 + (((c["D"] + 15) // 16 * 16) + 127) // 128 + 2
fold = os.environ.get("WMRL_TC_MODE") == "3"
tr = ws[:4 * njobs * 4 * 8].view(torch.int64).cpu().reshape(4, njobs, 4)
t0 = int(tr[1, 0, 0])
names = ["actor0", "actor1", "actor2", "mu", "embed"] + [f"gru{i}" for i in range(njobs - 7)] + ["prior1", "prior2"]
print("step 1 (cycles rel. to actor0 MMA start):  mma_start  mma_issued  epi_wake  epi_done | mma->wake  epi  done->next_mma")
for j in range(njobs):
    a, b, w_, e = [int(x) - t0 for x in tr[1, j]]
    nxt = int(tr[1, j + 1, 0]) - t0 if j + 1 < njobs else int(tr[2, 0, 0]) - t0
    print(f"  {names[j]:7s} {a:9d} {b:9d} {w_:9d} {e:9d} | {w_ - a:7d} {e - w_:7d} {nxt - e:7d}")
print("step length", int(tr[2, 0, 0]) - int(tr[1, 0, 0]))

pr = ws[600 * 8:(600 + 64) * 8].view(torch.int64).cpu()
def d(a, b): return int(pr[b]) - int(pr[a])
print("prior2 probes: math+stage", d(0, 1), " h'copy", d(1, 2), " barrier", d(2, 3), " copy-out", d(3, 4))
print("LN (last layer) probes: pass1", d(8, 9), " exchange+sync", d(9, 10), " pass2", d(10, 11))
print("job-end fences (tcgen05.fence + fence.proxy.async):", [d(16 + 2 * j, 17 + 2 * j) for j in range(njobs)])

mw = ws[700 * 8:704 * 8].view(torch.int64).cpu().tolist()
print(f"MMA-thread waits in step 1 (cycles): dep {mw[0]}  local full {mw[1]}  peer full {mw[2]}  over {mw[3]} stages")

x = ws[704 * 8:712 * 8].view(torch.int64).cpu().tolist()
print(f"actor1 (32 stages) per-stage avg: MMA lane wait full {x[0]/32:.0f}, wait peer {x[1]/32:.0f}, issue+commit {x[2]/32:.0f} | "
      f"producer wait empty {x[6]/32:.0f}, issue {x[7]/32:.0f}")

g = ws[712 * 8:716 * 8].view(torch.int64).cpu().tolist()
print(f"GRU stages ({g[3]}): total dep wait {g[0]}, per-stage avg wait full {g[1]/max(g[3],1):.0f}, issue+commit {g[2]/max(g[3],1):.0f}")

f = ws[720 * 8:725 * 8].view(torch.int64).cpu().tolist()
if f[4]:
    print(f"fold kernel, actor1 ({f[4]} stages) per-stage avg: wait full {f[0]/f[4]:.0f}, wait peer {f[1]/f[4]:.0f}, MMA issue {f[2]/f[4]:.0f}, commit {f[3]/f[4]:.0f}")

raw = ws[:1024 * 8].view(torch.int64).cpu()
print("dependency wait per job (step 1):", [int(raw[760 + j]) for j in range(njobs)])
print("traced job's stages (wait-weights, issue+commit):", [(int(raw[730 + 2 * i]), int(raw[731 + 2 * i])) for i in range(10)])
