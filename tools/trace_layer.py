"""Per-launch timeline of the layer-wise bf16 chain (csrc/tc_layer.cu), step 1, CTA 0 of every launch:
SM-clock offsets of producer start / first stage landed / first item's MMAs issued / all MMAs issued /
first epilogue wake + done / kernel end.  `python tools/trace_layer.py [c2|c5|c4]`."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "world-model-rl_b200"), os.path.join(ROOT, "tests")]
import torch
import helpers_gpu as HG
from wmrl import _lib, functional as Fn
from wmrl.rssm import _actor_tensors

case = (sys.argv[1:] or ["c2"])[0]
variant, D, S, Cc, Hd, A, Ha, La, B, H = {
    "c2": ("discrete", 600, 32, 32, 600, 6, 512, 3, 2500, 15),
    "c5": ("discrete", 1024, 32, 32, 1024, 6, 512, 3, 8192, 8),
    "c4": ("continuous", 200, 30, 1, 200, 6, 512, 3, 65536, 15),
}[case]
if case == "c4":
    os.environ["WMRL_BF16_PATH"] = "layer"
dev = torch.device("cuda")
rssm, actor = HG.build_modules(variant, D, S, Cc, Hd, 64, A, Ha, La, precision="bf16")
d0, s0, nz = [t.to(dev) for t in HG.synth_inputs(variant, B, H, D, S, Cc)]
lib = _lib.load()
La, Ha, actor_p = _actor_tensors(actor)
cfg = rssm._config(A, Ha, La, 5.0)
bound = torch.ones(A, device=dev)
rssm_p = rssm._rssm_tensors()
packed = rssm._packed.get(cfg, rssm_p, actor_p, bound)
S_in = S * Cc
deter = torch.empty(B, H + 1, D, device=dev); stoch = torch.empty(B, H + 1, S_in, device=dev)
da = torch.empty(B, H + 1, S_in, device=dev); db = torch.empty(B, H + 1, S_in, device=dev)
act = torch.empty(B, H, A, device=dev); idx = torch.empty(B, H, Cc, dtype=torch.int32, device=dev)
d = cfg.dims(B, H, 4, _lib.BF16)
ws_n = lib.wmrl_workspace_bytes(C.byref(d), 0)
ws = torch.zeros(ws_n, dtype=torch.uint8, device=dev)
w, aw = Fn._rssm_ptrs([p.detach() for p in rssm_p]), Fn._actor_ptrs([p.detach() for p in actor_p], La, bound)
for _ in range(2):
    _lib.check(lib.wmrl_imagine_fwd(C.byref(d), C.byref(w), C.byref(aw), packed.data_ptr(), d0.data_ptr(), s0.data_ptr(),
                                    nz.data_ptr(), deter.data_ptr(), stoch.data_ptr(), da.data_ptr(), db.data_ptr(),
                                    act.data_ptr(), idx.data_ptr(), None, 0, ws.data_ptr(), ws_n,
                                    torch.cuda.current_stream().cuda_stream), "fwd")
torch.cuda.synchronize()
nl = La + 5
tr = ws[:nl * 32 * 8].view(torch.int64).cpu().reshape(nl, 32)
names = [f"actor{l}" for l in range(La)] + ["mu", "embed", "gru", "prior1", "prior2"]
print(f"{case}: cycles relative to kernel start (CTA 0)")
print("  layer    setup  prod1st  prodItem0  stage0  mmaItem0  mmaAll  epiWait  epiWake  epi0Done  epiAll  end   items")
for j in range(nl):
    r = [int(x) for x in tr[j]]
    t0 = r[0]
    rel = lambda k: (r[k] - t0) if r[k] else -1
    print(f"  {names[j]:7s} {rel(1):7d} {rel(1):7d} {rel(2):9d} {rel(3):7d} {rel(4):8d} {rel(5):7d} {rel(6):8d} {rel(7):8d} {rel(8):8d} {rel(9):7d} {rel(11):6d}  {r[10]}")
pr = [int(x) for x in tr[nl - 1][16:22]]
if pr[5]:
    print("prior2 sample epilogue (thread 0, last block): tmem+bias", pr[1] - pr[0], " logits out", pr[2] - pr[1], " q in", pr[3] - pr[2],
          " softmax/argmax", pr[4] - pr[3], " one-hot out", pr[5] - pr[4])
