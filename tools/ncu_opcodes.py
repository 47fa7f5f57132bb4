#!/usr/bin/env python
"""Executed warp-instruction histogram by SASS opcode: python tools/ncu_opcodes.py X.ncu-rep [kernel-substring]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""
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
    hist = collections.Counter()
    for r in b["rows"]:
        toks = r[h["Source"]].split()
        op = next((t for t in toks if not t.startswith("@")), "?").split(".")[0]
        hist[op] += int(r[h["Instructions Executed"]])
    tot = sum(hist.values())
    print(f"== {b['name'][:70]}: {tot} warp instructions")
    print("   " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in hist.most_common(28)))
