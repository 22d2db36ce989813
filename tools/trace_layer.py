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

case = ([a for a in sys.argv[1:] if a != "bwd"] or ["c2"])[0]
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
def show(tr, names):
    print("  layer    setup  prod1st  prodItem0  stage0  mmaItem0  mmaAll  epiWait  epiWake  epi0Done  epiAll  end   items")
    for j in range(len(names)):
        r = [int(x) for x in tr[j]]
        t0 = r[0]
        rel = lambda k: (r[k] - t0) if r[k] else -1
        print(f"  {names[j]:9s} {rel(1):7d} {rel(1):7d} {rel(2):9d} {rel(3):7d} {rel(4):8d} {rel(5):7d} {rel(6):8d} {rel(7):8d} {rel(8):8d} {rel(9):7d} {rel(11):6d}  {r[10]}")


if "bwd" in sys.argv:
    # recording forward, then the BPTT with the trace flag: CTA 0 of every backward launch of step T-2
    d1 = cfg.dims(B, H, _lib.FLAG_SAVE_FOR_BWD, _lib.BF16)
    sv_n = lib.wmrl_saved_bytes(C.byref(d1), 0)
    saved = torch.zeros(sv_n, dtype=torch.uint8, device=dev)
    ws1_n = lib.wmrl_workspace_bytes(C.byref(d1), 0)
    ws1 = torch.zeros(ws1_n, dtype=torch.uint8, device=dev)
    _lib.check(lib.wmrl_imagine_fwd(C.byref(d1), C.byref(w), C.byref(aw), packed.data_ptr(), d0.data_ptr(), s0.data_ptr(),
                                    nz.data_ptr(), deter.data_ptr(), stoch.data_ptr(), da.data_ptr(), db.data_ptr(),
                                    act.data_ptr(), idx.data_ptr(), saved.data_ptr(), sv_n, ws1.data_ptr(), ws1_n,
                                    torch.cuda.current_stream().cuda_stream), "fwd(rec)")
    d2 = cfg.dims(B, H, _lib.FLAG_SAVE_FOR_BWD | _lib.FLAG_BACKWARD | 4, _lib.BF16)
    ws2_n = lib.wmrl_workspace_bytes(C.byref(d2), 0)
    ws2 = torch.zeros(ws2_n, dtype=torch.uint8, device=dev)
    ga = [torch.zeros_like(p) for p in actor_p]
    gaw = Fn._actor_ptrs(ga, La, None, grads=True)
    g_de = torch.randn(B, H + 1, D, device=dev) / B
    g_st = torch.randn(B, H + 1, S_in, device=dev) / B
    g_d0 = torch.zeros(B, D, device=dev); g_s0 = torch.zeros(B, S_in, device=dev)
    for _ in range(2):
        _lib.check(lib.wmrl_imagine_bwd(C.byref(d2), C.byref(w), C.byref(aw), packed.data_ptr(), nz.data_ptr(), deter.data_ptr(),
                                        stoch.data_ptr(), da.data_ptr(), db.data_ptr(), act.data_ptr(), saved.data_ptr(),
                                        g_de.data_ptr(), g_st.data_ptr(), None, None, C.byref(gaw), None, g_d0.data_ptr(),
                                        g_s0.data_ptr(), ws2.data_ptr(), ws2_n, torch.cuda.current_stream().cuda_stream), "bwd")
    torch.cuda.synchronize()
    nb = 5 + La
    trb = ws2[:nb * 32 * 8].view(torch.int64).cpu().reshape(nb, 32)
    print(f"{case} backward (step T-2): cycles relative to kernel start (CTA 0)")
    show(trb, ["prior2^T", "prior1^T+GRUb", "W_ih^T", "W_hh^T", "embed^T"] + [f"LNb{l}" for l in range(La - 1, -1, -1)])
    pc = [int(x) for x in trb[3][16:20]]
    print("W_hh^T carry epilogue probes (thread 0, last block): loads issued+staged", pc[1] - pc[0], " tmem wait", pc[2] - pc[1], " math+stores", pc[3] - pc[2])
    pl = [int(x) for x in trb[nb - 1][16:21]]
    print("LNb0 probes: first block loads", pl[1] - pl[0], " pass1", pl[2] - pl[0], " exchange", pl[3] - pl[2], " pass2", pl[4] - pl[3])

pr = [int(x) for x in tr[nl - 1][16:22]]
if pr[5]:
    print("prior2 sample epilogue (thread 0, last block): tmem+bias", pr[1] - pr[0], " logits out", pr[2] - pr[1], " q in", pr[3] - pr[2],
          " softmax/argmax", pr[4] - pr[3], " one-hot out", pr[5] - pr[4])
