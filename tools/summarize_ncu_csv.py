"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (second half = the warm pass).
   python tools/summarize_ncu_csv.py gpurun_out/x.csv [n_passes]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
data = rows[hdr + 1:]
data = data[len(data) - len(data) // passes:]
agg = collections.OrderedDict()
for r in data:
    n = r[ki].split("(")[0].replace("dlb::", "").replace("void ", "")
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:70]:70s} n={a[0]:4d} us={a[1]:10.1f} share={a[1] / tot:.3f}")
print("total ms", tot / 1e3, "launches", len(data))
