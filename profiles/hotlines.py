"""Stall breakdown and hottest source lines of a kernel from an ncu report captured with --import-source on:
python profiles/hotlines.py <report.ncu-rep> [top_n]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr = None; fname = ""
lines = []
stalls = collections.Counter()
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 10 or r[0] == "":
        continue            # SASS rows (empty line number) are already summed into their CUDA line
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)
    try:
        smp = int(d["# Samples"]); ins = int(d["Instructions Executed"])
    except (ValueError, KeyError):
        continue
    lines.append((fname, r[0], r[1].strip()[:105], smp, ins))
    for k, v in zip(hdr, r):
        if k.startswith("stall_") and "Not Issued" not in k:
            try: stalls[k] += int(v or 0)
            except ValueError: pass
tot_s = sum(l[3] for l in lines) or 1; tot_i = sum(l[4] for l in lines) or 1
print("total samples", tot_s, "warp instructions", tot_i)
ts = sum(stalls.values()) or 1
print({k: round(100.0 * v / ts, 1) for k, v in stalls.most_common(9)})
for f, ln, src, smp, ins in sorted(lines, key=lambda l: -l[3])[:topn]:
    print(f"{f}:{ln:>4} samp {100.0*smp/tot_s:5.1f}% inst {100.0*ins/tot_i:5.1f}%  {src}")
