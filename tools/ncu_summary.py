"""Summarise an .ncu-rep: key raw metrics + per-opcode / per-stall aggregation of the source page.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--hot N]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
hot = int(sys.argv[sys.argv.index("--hot") + 1]) if "--hot" in sys.argv else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu", "gpc__cycles_elapsed.avg.per_second"]
for h, u, v in zip(hdr, units, vals):
    if any(h == w or (w.endswith("fma") and h.startswith(w)) or (w.endswith("alu") and h.startswith(w)) for w in want):
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S, E = ix["# Samples"], ix["Instructions Executed"]
print("total samples", sum(int(r[S]) for r in data), "warp-inst", sum(int(r[E]) for r in data))
def opname(t):
    t = t.strip().split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
byop, bys = collections.Counter(), collections.Counter()
for r in data:
    byop[opname(r[1])] += int(r[E]); bys[opname(r[1])] += int(r[S])
print("opcode: executed(M) samples")
for op, c in byop.most_common(22):
    print(f"  {op:10s} {c/1e6:9.1f} {bys[op]:8d}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    for h in stalls:
        tot[h] += int(r[ix[h]])
print("stalls:", tot.most_common(9))
print("hottest SASS (addr samples executed text)")
for r in sorted(data, key=lambda r: -int(r[S]))[:hot]:
    print("  ", r[0][-5:], r[S].rjust(7), r[E].rjust(10), r[1][:80])
