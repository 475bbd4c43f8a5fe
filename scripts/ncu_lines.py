"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line:
   python scripts/ncu_lines.py <dump.csv> [top]   -> file:line, executed warp instructions, stall samples, source"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = []
fname = ""
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if r[0] == "Function Name":
        continue
    if r[0] and r[0].isdigit() and hdr:
        try:
            samples = int(r[6]); inst = int(r[7]); thr = int(r[8])
        except ValueError:
            continue
        out.append((fname, int(r[0]), inst, samples, thr, r[1].strip()))
ti = sum(o[2] for o in out); ts = sum(o[3] for o in out)
print(f"total warp inst {ti}  samples {ts}")
print("--- by samples")
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{o[0]}:{o[1]:<4} inst {o[2]:>9} ({100*o[2]/ti:4.1f}%) samples {o[3]:>6} ({100*o[3]/ts:4.1f}%) lanes {o[4]/max(o[2],1):4.1f} | {o[5][:90]}")
