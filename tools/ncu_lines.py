#!/usr/bin/env python
"""Executed warp instructions and stall samples per SOURCE line: python tools/ncu_lines.py X.ncu-rep kernel-substring [top]
(reads `ncu -i X --page source --csv --print-source cuda,sass`; needs -lineinfo and --import-source on)"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; pat = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fn, path, hdr = None, None, None
agg = collections.defaultdict(lambda: [0, 0, ""])
done = set()
for r in rows:
    if not r: continue
    if r[0] == "File Path": path = r[1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = {k: i for i, k in enumerate(r)}; continue
    if hdr is None or fn is None or pat not in fn: continue
    if r[0] == "": continue
    # a source-line row: (line, text, '-', '-', samples..., instructions executed)
    try:
        key = (fn, path.split("/")[-1], int(r[0]))
        a = agg[key]
        a[0] += int(r[hdr["Instructions Executed"]]); a[1] += int(r[hdr["# Samples"]]); a[2] = r[1].strip()
    except Exception:
        pass
fns = sorted({k[0] for k in agg})
for f in fns[:1]:
    items = [(k, v) for k, v in agg.items() if k[0] == f]
    ti = sum(v[0] for _, v in items); ts = sum(v[1] for _, v in items)
    print(f"== {f[:90]}: {ti} warp instructions, {ts} stall samples")
    for k, v in sorted(items, key=lambda kv: -kv[1][0])[:top]:
        print(f"  {k[1]:18s}:{k[2]:4d} inst {100 * v[0] / max(ti, 1):5.1f}%  stall {100 * v[1] / max(ts, 1):5.1f}%  {v[2][:110]}")
# optional: sums over line ranges of one file: MV_RANGES="k_oit.cu:154-177,k_oit.cu:82-136"
import os
if os.environ.get("MV_RANGES"):
    f = fns[0]
    ti = sum(v[0] for k, v in agg.items() if k[0] == f)
    for spec in os.environ["MV_RANGES"].split(","):
        name, rng = spec.split(":"); a, b = map(int, rng.split("-"))
        tot = sum(v[0] for k, v in agg.items() if k[0] == f and k[1] == name and a <= k[2] <= b)
        print(f"  range {spec:28s} {100 * tot / ti:5.1f}% of instructions")
