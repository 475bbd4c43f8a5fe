"""Top SASS instructions of a kernel by warp-stall samples, from `ncu -i rep --page source --csv` (run here).
   python scripts/ncu_lines.py <source-page csv> [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
col = {n: i for i, n in enumerate(h)}
data = rows[hi + 1:]
tot = sum(int(r[col["# Samples"]]) for r in data if len(r) > col["# Samples"])
inst = sum(int(r[col["Instructions Executed"]]) for r in data)
print("total samples", tot, "warp instructions", inst)
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = {s: sum(int(r[col[s]]) for r in data) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
order = sorted(range(len(data)), key=lambda k: -int(data[k][col["# Samples"]]))[:top]
for k in sorted(order):
    r = data[k]
    why = sorted(((int(r[col[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{k:5d} {int(r[col['# Samples']]):6d} {100*int(r[col['# Samples']])/tot:5.1f}%  exec {int(r[col['Instructions Executed']]):8d} thr {r[col['Avg. Threads Executed']]:>5s}  {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}   {r[col['Source']].strip()}")
