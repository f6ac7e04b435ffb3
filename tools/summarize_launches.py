"""Summarise an `ncu --metrics gpu__time_duration.sum` launch list (CSV) of `bench.py --profile-eager`:
per-kernel share of ONE training step (the last of the profiled steps).  Usage:
  python tools/summarize_launches.py gpurun_out/launches_r01.csv 4 > profiles/launches_r01_summary.md"""
import collections, csv, re, sys

path, nsteps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 4
lines = [l for l in open(path) if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    try:
        t = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    if row["Metric Unit"] == "ns":
        t /= 1e3
    name = re.sub(r"<.*", "", row["Kernel Name"].split("(")[0]).replace("void ", "").replace("dlb::", "")
    rows.append((name, t))
per = len(rows) // nsteps
last = rows[-per:]
tot = sum(t for _, t in last)
agg = collections.defaultdict(lambda: [0, 0.0])
for n, t in last:
    agg[n][0] += 1
    agg[n][1] += t
print(f"# ncu launch list, one training step (bs 16, 512x512, fp16): {per} launches, {tot / 1e3:.2f} ms serialised (cold cache)\n")
print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
