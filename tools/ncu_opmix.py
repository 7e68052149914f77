#!/usr/bin/env python3
"""Summarises an `ncu --page source --csv` export: executed warp instructions by opcode and stall reasons.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_opmix.py src.csv [kernel-index]"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
# split per kernel: a "Kernel Name" row followed by a header row
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["data"].append(r)
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
k = kernels[which]
hdr = k["hdr"]; ci = {h: i for i, h in enumerate(hdr)}
tot = collections.Counter(); samp = collections.Counter(); thr = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stall = collections.Counter()
for r in k["data"]:
    if len(r) < len(hdr):
        continue
    sass = r[ci["Source"]].strip()
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", sass)
    op = m.group(2) if m else sass
    n = int(r[ci["Instructions Executed"]]); s = int(r[ci["# Samples"]])
    tot[op] += n; samp[op] += s; thr[op] += int(r[ci["Thread Instructions Executed"]])
    for h in stall_cols:
        stall[h] += int(r[ci[h]] or 0)
ti, ts = sum(tot.values()), sum(samp.values())
print(k["name"][:100], f"({len(kernels)} kernels in file)")
print("warp instructions executed:", ti, " stall samples:", ts, " avg active threads:", round(sum(thr.values()) / ti, 2))
for op, n in tot.most_common(36):
    print(f"  {op:10s} {n:13d} {n / ti:6.1%}   samples {samp[op] / ts:6.1%}")
print("stalls:", {kk.replace("stall_", ""): f"{v / ts:.1%}" for kk, v in stall.most_common(9)})
