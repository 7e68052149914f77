#!/usr/bin/env python3
"""Instruction working set of a kernel from an ncu capture: how many SASS instructions (and 128-byte cache lines) account for the
executed instructions.  The BVH kernels are issue-bound with divergent warps all over the loop body; once the lines that cover
99 % of the executed instructions exceed the 32 KB L1.5 instruction cache, `no_instruction` stalls take over (profiles/README.md).

    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_hotset.py src.csv
"""
import csv, sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[0].startswith("0x")]
ex = [int(r[ci["Instructions Executed"]]) for r in data]
mx, tot = max(ex), sum(ex)
print("instructions", len(ex), "max executions of one instruction", mx)
for frac in (0.5, 0.2, 0.1, 0.05, 0.01, 0.001):
    hot = [e for e in ex if e >= frac * mx]
    print(f"  executed >= {frac:5.3f} of the maximum: {len(hot):5d} instructions = {len(hot) * 16 / 1024:5.1f} KB, covering {sum(hot) / tot:6.1%} of all executed")
lines = {}
for i, e in enumerate(ex):
    lines[i // 8] = lines.get(i // 8, 0) + e
cum = 0
for k, v in enumerate(sorted(lines.values(), reverse=True)):
    cum += v
    if cum >= 0.99 * tot:
        print("  128-byte lines covering 99 % of the executed instructions:", k + 1, "=", (k + 1) * 128 / 1024, "KB")
        break
