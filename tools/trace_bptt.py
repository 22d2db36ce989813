"""Cycle timeline of the persistent BPTT kernel (WMRL_FLAG_TRACE): per backward step and job of CTA 0's first
tile - when the MMA lane issued the job's first MMA and committed its accumulator, when the epilogue warps woke
and finished.  Drives the C ABI directly.

    python tools/trace_bptt.py [--B 65536] [--H 15]
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "world-model-rl_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

from wmrl import _lib  # noqa: E402
from wmrl import functional as WF  # noqa: E402

JOBS = ["prior2^T (ELU')", "prior1^T + GRU bwd", "W_ih^T (ELU')", "W_hh^T carry", "embed^T + tanh + sample bwd",
        "mu^T + LN2 bwd", "W2^T + LN1 bwd", "W1^T + LN0 bwd"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=65536)
    ap.add_argument("--H", type=int, default=15)
    a = ap.parse_args()
    import helpers_gpu as HG
    from wmrl.rssm import _actor_bound, _actor_tensors
    D, S, Hd, A, Ha, La = 200, 30, 200, 6, 512, 3
    B, H = a.B, a.H
    dev = torch.device("cuda")
    rssm, actor = HG.build_modules("continuous", D, S, 1, Hd, 64, A, Ha, La, seed=0, precision="bf16")
    d0, s0, nz = (x.to(dev) for x in HG.synth_inputs("continuous", B, H, D, S, 1))
    lib = _lib.load()
    cfg = rssm._config(A, Ha, La, actor.action_scale)
    _, _, actor_p = _actor_tensors(actor)
    rssm_p = rssm._rssm_tensors()
    bound = _actor_bound(actor, A, dev)
    packed = rssm._packed.get(cfg, rssm_p, actor_p, bound)
    d = cfg.dims(B, H, _lib.FLAG_SAVE_FOR_BWD, _lib.BF16)
    saved_n = lib.wmrl_saved_bytes(C.byref(d), 0)
    saved = torch.empty(saved_n, dtype=torch.uint8, device=dev)
    ws_n = lib.wmrl_workspace_bytes(C.byref(d), 0)
    ws = torch.empty(ws_n, dtype=torch.uint8, device=dev)
    a32 = [p.detach().float().contiguous() for p in actor_p]
    r32 = [p.detach().float().contiguous() for p in rssm_p]
    T1 = H + 1
    deter_seq, stoch_seq = torch.empty(B, T1, D, device=dev), torch.empty(B, T1, S, device=dev)
    means, stds, actions = torch.empty(B, T1, S, device=dev), torch.empty(B, T1, S, device=dev), torch.empty(B, H, A, device=dev)
    w, aw = WF._rssm_ptrs(r32), WF._actor_ptrs(a32, La, bound)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.wmrl_imagine_fwd(C.byref(d), C.byref(w), C.byref(aw), packed.data_ptr(), d0.data_ptr(), s0.data_ptr(),
                                    nz.contiguous().data_ptr(), deter_seq.data_ptr(), stoch_seq.data_ptr(), means.data_ptr(),
                                    stds.data_ptr(), actions.data_ptr(), None, saved.data_ptr(), saved_n, ws.data_ptr(), ws_n,
                                    st), "fwd")
    db = cfg.dims(B, H, _lib.FLAG_SAVE_FOR_BWD | _lib.FLAG_BACKWARD | _lib.FLAG_TRACE, _lib.BF16)
    wsb_n = lib.wmrl_workspace_bytes(C.byref(db), 0)
    wsb = torch.zeros(wsb_n, dtype=torch.uint8, device=dev)
    ga = [torch.zeros_like(p) for p in a32]
    gaw = WF._actor_ptrs(ga, La, None, grads=True)
    g_d0, g_s0 = torch.zeros(B, D, device=dev), torch.zeros(B, S, device=dev)
    gd = torch.randn(B, T1, D, device=dev) / B
    gs = torch.randn(B, T1, S, device=dev) / B
    for _ in range(2):
        _lib.check(lib.wmrl_imagine_bwd(C.byref(db), C.byref(w), C.byref(aw), packed.data_ptr(), nz.contiguous().data_ptr(),
                                        deter_seq.data_ptr(), stoch_seq.data_ptr(), means.data_ptr(), stds.data_ptr(),
                                        actions.data_ptr(), saved.data_ptr(), gd.data_ptr(), gs.data_ptr(), None, None,
                                        C.byref(gaw), None, g_d0.data_ptr(), g_s0.data_ptr(), wsb.data_ptr(), wsb_n, st), "bwd")
    torch.cuda.synchronize()
    tr = wsb[-65536 - 256:].cpu()
    # the trace sits right behind the dL/dh scratch: locate it from the layout used by tc_imagine_bwd
    ntiles = (B + 127) // 128
    Dp = (D + 15) // 16 * 16
    nblk = ntiles * 2 * H
    from bptt_debug import stash_layout
    _, _, _, blk_g = stash_layout(Dp, (S + 15) // 16 * 16, Ha, (Hd + 15) // 16 * 16, La)
    off = (nblk * blk_g + 255) // 256 * 256 + ntiles * 128 * Dp * 4
    t = wsb[off: off + 8 * 4 * 8 * 4].view(torch.int64).cpu().view(4, 8, 4)
    njobs = 5 + La
    for ti in range(1, 3):
        base = int(t[ti, 0, 0])
        print(f"backward step {ti} (cycles relative to the step's first MMA; step length "
              f"{int(t[ti + 1, 0, 0]) - base if ti + 1 < 4 else -1})")
        print(f"  {'job':30s} {'mma first':>10s} {'mma commit':>11s} {'epi wake':>10s} {'epi done':>10s} {'mma':>8s} {'epi':>8s}")
        for j in range(njobs):
            m0, m1, e0, e1 = (int(x) - base for x in t[ti, j])
            print(f"  {JOBS[j]:30s} {m0:10d} {m1:11d} {e0:10d} {e1:10d} {m1 - m0:8d} {e1 - e0:8d}")


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    main()
