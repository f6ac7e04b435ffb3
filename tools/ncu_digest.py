"""Digest of one `ncu --set full --import-source on` report: headline metrics + instructions / stall samples per source
line.  Usage: python tools/ncu_digest.py gpurun_out/x.ncu-rep [top_n] [--sass]   (ncu must be on PATH; no GPU needed).
--sass appends the executed-instruction histogram by SASS opcode and the most-sampled SASS instructions (what found the
per-thread tile decode and the CAS-loop shared atomics in round 1)."""
import csv, io, subprocess, sys

args = [a for a in sys.argv[1:] if not a.startswith("--")]
rep = args[0]
top = int(args[1]) if len(args) > 1 else 25
want_sass = "--sass" in sys.argv


def run(*a):
    return subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout


import os
if not os.path.exists(rep):
    sys.exit(f"{rep}: no such report (gpurun_out/ is scratch -- re-run profiles/run_profile.sh)")
rows = list(csv.reader(io.StringIO(run("--page", "raw", "--csv"))))
if len(rows) < 3:
    sys.exit(f"{rep}: ncu printed no kernel rows")
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
print("## metrics")
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h} [{u}] = {v}")
print("## stall reasons (warps per issue-active cycle)")
for h, u, v in zip(hdr, units, vals):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
        try:
            if float(v) >= 0.1:
                print(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:22s} {float(v):.2f}")
        except ValueError:
            pass

rows = list(csv.reader(io.StringIO(run("--page", "source", "--csv", "--print-source", "cuda,sass"))))
cur, hd, agg = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hd = {h: i for i, h in enumerate(r)}; continue
    if r[0] == "Function Name" or hd is None or len(r) < 8:
        continue
    if r[2] == "-":
        try:
            n = float(r[hd["Instructions Executed"]]); s = float(r[hd["# Samples"]])
        except ValueError:
            continue
        if n or s:
            agg[(cur, int(r[0]), r[1].strip()[:100])] = (n, s)
tn = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
print(f"## per source line: {tn:.0f} warp instructions, {ts:.0f} samples")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0] - kv[1][1] * tn / ts)[:top]:
    print(f"{k[0]:>14s}:{k[1]:<5d} inst {100 * v[0] / tn:5.1f}%  samples {100 * v[1] / ts:5.1f}%  {k[2]}")

if want_sass:
    import collections
    rows = list(csv.reader(io.StringIO(run("--page", "source", "--csv", "--print-source", "sass"))))
    hdr = next((r for r in rows if r and r[0] == "Address"), None)
    if hdr:
        ix = {h: i for i, h in enumerate(hdr)}
        byop, samp, recs, tot = collections.Counter(), collections.Counter(), [], 0
        for r in rows:
            if len(r) <= ix["Instructions Executed"] or r[0] == "Address":
                continue
            try:
                n = int(r[ix["Instructions Executed"]]); sm = int(r[ix["# Samples"]])
            except ValueError:
                continue
            toks = r[1].strip().split()
            if not toks:
                continue
            op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
            byop[op] += n; samp[op] += sm; tot += n
            recs.append((n, sm, r[1].strip()))
        ts = sum(samp.values()) or 1
        print(f"## SASS opcodes: {tot} warp instructions, {ts} samples")
        for op, n in byop.most_common(top):
            print(f"{op:12s} inst {100 * n / max(tot, 1):5.1f}%  samples {100 * samp[op] / ts:5.1f}%")
        print("## most-sampled SASS instructions")
        for n, sm, t in sorted(recs, key=lambda x: -x[1])[:top]:
            print(f"samples {100 * sm / ts:5.1f}%  executed {n:9d}  {t[:100]}")
