#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv`:
python tools/ncu_top.py X.ncu-rep [kernel-substring] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(txt)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": [], "hdr": None}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
seen = set()
for b in blocks:
    if pat not in b["name"] or b["name"] in seen:
        continue
    seen.add(b["name"])
    h = {k: i for i, k in enumerate(b["hdr"])}
    rows = b["rows"]
    tot = sum(int(r[h["# Samples"]]) for r in rows); inst = sum(int(r[h["Instructions Executed"]]) for r in rows)
    thr = sum(int(r[h["Thread Instructions Executed"]]) for r in rows)
    print(f"== {b['name'][:70]}: stall samples {tot}, warp instructions {inst}, avg threads/inst {thr / max(inst, 1):.1f}")
    reasons = [k for k in b["hdr"] if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(int(r[h[k]]) for r in rows) for k in reasons}
    print("   stall reasons:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(rows, key=lambda r: -int(r[h["# Samples"]]))[:top]:
        s = int(r[h["# Samples"]])
        print(f"   {r[h['Address']][-5:]} {s:7d} {100 * s / max(tot, 1):5.1f}%  long_sb {int(r[h['stall_long_sb']]):6d}  exec {int(r[h['Instructions Executed']]):9d}  {r[h['Source']].strip()}")
