"""Warp-stall samples per SOURCE line for one kernel: joins the per-instruction samples of
`ncu -i rep --page source --csv --print-source=sass --kernel-name regex:K` with the line table of
`nvdisasm -g -c <cubin>` (same instruction order).
    python tools/sass_lines.py samples.csv disasm.txt <mangled-name substring> [top N]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
idx = {n: i for i, n in enumerate(h)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(h) and r[idx["# Samples"]].isdigit():
        data.append(r)
txt = open(sys.argv[2]).read().split("\n")
key = sys.argv[3]
N = int(sys.argv[4]) if len(sys.argv) > 4 else 40
start = next(i for i, l in enumerate(txt) if l.lstrip().startswith(".section") and ".text." in l and key in l)
ins, cur = [], None
for l in txt[start + 1:]:
    if l.lstrip().startswith(".section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        cur = (m.group(1).split("/")[-1], int(m.group(2)), tuple((f.split("/")[-1], int(n)) for f, n in inl))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((cur, m.group(2)))
print("instructions: disasm", len(ins), "profile", len(data))
if len(ins) != len(data):
    sys.exit("instruction counts differ: the cubin is not the one that was profiled")
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.Counter()
why = collections.defaultdict(collections.Counter)
for (cur, _), r in zip(ins, data):
    n = int(r[idx["# Samples"]])
    k = (cur[0], cur[1], cur[2][-1] if cur[2] else None) if cur else None
    agg[k] += n
    for s in stalls:
        why[k][s[6:]] += int(r[idx[s]])
tot = sum(agg.values())
print("total samples", tot)
for k, v in agg.most_common(N):
    print(f"{100 * v / tot:5.1f}%  {v:6d}  {k}  {why[k].most_common(3)}")
