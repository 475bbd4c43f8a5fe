"""Summarise ncu captures for profiles/ (run here, after gpurun brought the files back).
   python scripts/ncu_summary.py launches <launch list csv>            -> per-kernel share table (markdown)
   python scripts/ncu_summary.py full <report.ncu-rep> [...]           -> per-kernel metric table (markdown) + traffic json
"""
import csv, io, json, subprocess, sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("smsp__inst_executed.sum", "warp instr"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/instr"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %")]
STALLS = ["long_scoreboard", "wait", "short_scoreboard", "lg_throttle", "mio_throttle", "branch_resolving", "not_selected",
          "no_instruction", "math_pipe_throttle", "barrier"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = {}
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", "")) / 1000
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | mean us | sum us | share |\n|---|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{n[:60]}` | {c} | {t / c:.1f} | {t:.1f} | {100 * t / tot:.1f}% |")
    print(f"\ntotal {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")


def full(paths):
    seen, traffic = {}, {}
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        r = list(csv.reader(io.StringIO(out)))
        h, units = r[0], r[1]
        for row in r[2:]:
            name = row[h.index("Kernel Name")].split("(")[0].replace("void ", "")
            if name in seen:
                continue
            d = {}
            for k, label in WANT:
                if k in h:
                    v, u = row[h.index(k)], units[h.index(k)]
                    try:
                        v = float(v)
                    except ValueError:
                        pass
                    d[label] = (v, u)
            for s in STALLS:
                k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
                if k in h:
                    d["stall " + s] = (float(row[h.index(k)]), "")
            seen[name] = d

            def to_bytes(x):
                v, u = x
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            traffic[name] = int(to_bytes(d["DRAM read"]) + to_bytes(d["DRAM write"]))
    labels = [l for _, l in WANT]
    print("| kernel | " + " | ".join(labels) + " |\n|" + "---|" * (len(labels) + 1))
    for n, d in seen.items():
        cells = []
        for l in labels:
            v, u = d.get(l, ("", ""))
            cells.append(f"{v:.1f} {u}".strip() if isinstance(v, float) else str(v))
        print(f"| `{n}` | " + " | ".join(cells) + " |")
    print("\nstall reasons (warps stalled per issued instruction):\n")
    print("| kernel | " + " | ".join(STALLS) + " |\n|" + "---|" * (len(STALLS) + 1))
    for n, d in seen.items():
        print(f"| `{n}` | " + " | ".join(f"{d.get('stall ' + s, (0, ''))[0]:.2f}" for s in STALLS) + " |")
    print("\nTRAFFIC_JSON " + json.dumps(traffic))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
