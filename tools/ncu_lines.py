#!/usr/bin/env python3
"""Attributes the executed warp instructions of an ncu capture to device functions / source lines.

  cuobjdump -xelf all rt_b200/lib/librtcu.so && nvdisasm --print-line-info-inline rtcu.sm_100a.cubin > dis.txt
  ncu -i rep.ncu-rep --page source --csv > src.csv
  python tools/ncu_lines.py src.csv dis.txt <mangled-kernel-substring> [--lines] [--kernel=N]   (N-th kernel section of the csv)

Every SASS instruction is charged to the innermost source line nvdisasm reports for it (through inlining) and,
from there, to the enclosing __device__/__global__ function found by scanning the source file.
"""
import collections, csv, re, sys

src_csv, dis_txt, kernel_sub = sys.argv[1:4]
by_line = "--lines" in sys.argv

# ---- executed counts per instruction offset from ncu
rows = list(csv.reader(open(src_csv)))
which = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--kernel=")), 0)
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if starts:
    lo = starts[which]
    hi = starts[which + 1] if which + 1 < len(starts) else len(rows)
    print(rows[lo][1][:110])
    rows = rows[lo:hi]
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) >= len(hdr) and r[0].startswith("0x")]
base = int(data[0][0], 16)
execd = {int(r[0], 16) - base: (int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]]), r[ci["Source"]].strip()) for r in data}
threads = {int(r[0], 16) - base: int(r[ci["Thread Instructions Executed"]]) for r in data} if "Thread Instructions Executed" in ci else {}

# ---- line info per offset from nvdisasm
lines = open(dis_txt).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kernel_sub in l and l.rstrip().endswith(":"))
info, cur = {}, []
for l in lines[start + 1:]:
    if l.startswith("//---") or l.startswith("\t.section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur.append((m.group(1), int(m.group(2)), "inlined at" in m.group(3)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", l)
    if m:
        off = int(m.group(1), 16)
        if cur:
            info[off] = cur[0][:2]  # first annotation = innermost location
            last = cur[0][:2]
            cur = []
        else:
            info[off] = info.get(off, None) or last_seen if (last_seen := locals().get("last")) else None
        last = info[off] if info[off] else locals().get("last")

# ---- function ranges by scanning sources
def function_ranges(path):
    out = []
    try:
        text = open(path).read().split("\n")
    except OSError:
        return out
    for i, l in enumerate(text):
        m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__)[^;(]*?\b(\w+)\s*\(", l)
        if m:
            out.append((i + 1, m.group(1)))
    return out

ranges = {}
def func_of(path, line):
    if path not in ranges:
        ranges[path] = function_ranges(path)
    name = "?"
    for ln, fn in ranges[path]:
        if ln <= line:
            name = fn
        else:
            break
    return name

agg = collections.Counter(); samp = collections.Counter(); thr = collections.Counter()
total = 0; tsamp = 0
for off, (n, s, sass) in execd.items():
    loc = info.get(off)
    if loc is None:
        key = "(no line info)"
    else:
        path, line = loc
        short = path.split("/")[-1]
        key = f"{short}:{line}" if by_line else f"{short}:{func_of(path, line)}"
    agg[key] += n; samp[key] += s; total += n; tsamp += s; thr[key] += threads.get(off, 0)
print(f"total warp instructions {total}, stall samples {tsamp}, active lanes per warp instruction {sum(thr.values()) / max(1, total):.2f}")
for k, n in agg.most_common(45 if by_line else 30):
    print(f"  {k:45s} {n:13d} {n / total:6.1%}   samples {samp[k] / max(1, tsamp):6.1%}   lanes {thr[k] / max(1, n):5.1f}")
