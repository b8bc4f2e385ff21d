"""Hot instructions of one kernel from `ncu -i rep --page source --csv --print-source=sass` output.
    python tools/sass_hot.py file.csv [top N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = rows[1]
idx = {n: i for i, n in enumerate(h)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(h) and r[idx["# Samples"]].isdigit():
        data.append(r)
tot = sum(int(r[idx["# Samples"]]) for r in data)
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = {s: sum(int(r[idx[s]]) for r in data) for s in stalls}
print("kernel", rows[0][1][:80], "| samples", tot, "| instructions", len(data))
print("stall totals:", [(k[6:], v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]])
top = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))[:N]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[idx[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print(str(i).rjust(6), r[idx["# Samples"]].rjust(6), r[idx["Instructions Executed"]].rjust(9), r[idx["Source"]].strip()[:72].ljust(72), st)
