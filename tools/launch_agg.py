"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; data = rows[hdr + 1:]
last = int(sys.argv[2]) if len(sys.argv) > 2 else len(data)
ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for r in data[-last:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    us = v / 1e3 if u == "ns" else (v if u == "us" else v * 1e3)
    k = r[ki].split("(")[0][-44:]
    agg[k][0] += 1; agg[k][1] += us; tot += us
print(f"launches {sum(n for n, _ in agg.values())}  total {tot:.1f} us")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:10.1f} us {100*us/tot:5.1f}%  x{n:4d}  avg {us/n:8.1f} us  {k}")
