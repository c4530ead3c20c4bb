"""Executed warp instructions per source line (>= min_pct of the kernel), in file/line order, from an ncu report with
--import-source on:  python profiles/lines_by_number.py <report.ncu-rep> [min_pct]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr = None; fname = ""; rows = []
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 10 or r[0] == "":
        continue
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)
    try:
        rows.append((fname, int(r[0]), r[1].strip()[:100], float(d["Instructions Executed"] or 0), float(d["Thread Instructions Executed"] or 0), float(d["# Samples"] or 0)))
    except (ValueError, KeyError):
        pass
T = sum(r[3] for r in rows) or 1.0; S = sum(r[5] for r in rows) or 1.0
print(f"total warp instructions {T:.4g}")
cum = 0.0
for f, ln, src, ins, tins, smp in sorted(rows, key=lambda r: (r[0], r[1])):
    if 100 * ins / T >= minp:
        print(f"{f}:{ln:>4} inst {100*ins/T:5.2f}% samp {100*smp/S:5.2f}% lanes {tins/max(ins,1):4.1f}  {src}")
