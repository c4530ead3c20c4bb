"""Per-source-line instruction, sample, shared-memory-wavefront and L2-sector totals of a kernel from an ncu report
captured with --import-source on:  python profiles/regions.py <report.ncu-rep> [top_n]
(the SASS rows of a CUDA line are already summed into it by ncu)"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
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
    def num(k):
        try: return float(d.get(k, 0) or 0)
        except ValueError: return 0.0
    rows.append(dict(file=fname, line=int(r[0]), src=r[1].strip()[:90], samp=num("# Samples"), inst=num("Instructions Executed"),
                     tinst=num("Thread Instructions Executed"), shw=num("L1 Wavefronts Shared"), shi=num("L1 Wavefronts Shared Ideal"),
                     l2=num("L2 Theoretical Sectors Global"), l2i=num("L2 Theoretical Sectors Global Ideal"), tag=num("L1 Tag Requests Global")))
T = {k: sum(r[k] for r in rows) or 1.0 for k in ("samp", "inst", "tinst", "shw", "shi", "l2", "l2i", "tag")}
print(f"warp instructions {T['inst']:.4g}  thread instructions {T['tinst']:.4g} (avg {T['tinst']/T['inst']:.1f} active lanes)  samples {T['samp']:.4g}")
print(f"shared-memory wavefronts {T['shw']:.4g} (ideal {T['shi']:.4g}, x{T['shw']/T['shi']:.2f})   L2 sectors (theoretical) {T['l2']:.4g} (ideal {T['l2i']:.4g})   L1 tag requests {T['tag']:.4g}")
byfile = collections.defaultdict(lambda: collections.Counter())
for r in rows:
    for k in ("samp", "inst", "shw", "l2"):
        byfile[r["file"]][k] += r[k]
for f, c in byfile.items():
    print(f"  {f:22s} inst {100*c['inst']/T['inst']:5.1f}%  samples {100*c['samp']/T['samp']:5.1f}%  shared wavefronts {100*c['shw']/T['shw']:5.1f}%  L2 sectors {100*c['l2']/T['l2']:5.1f}%")
print("# top lines by executed instructions")
for r in sorted(rows, key=lambda r: -r["inst"])[:topn]:
    print(f"{r['file']}:{r['line']:>4} inst {100*r['inst']/T['inst']:5.1f}% samp {100*r['samp']/T['samp']:5.1f}% lanes {r['tinst']/max(r['inst'],1):4.1f}  {r['src']}")
print("# top lines by shared-memory wavefronts (actual / ideal)")
for r in sorted(rows, key=lambda r: -r["shw"])[:12]:
    print(f"{r['file']}:{r['line']:>4} wavefronts {100*r['shw']/T['shw']:5.1f}%  x{r['shw']/max(r['shi'],1):.2f} of ideal  {r['src']}")
print("# top lines by L2 sectors")
for r in sorted(rows, key=lambda r: -r["l2"])[:12]:
    print(f"{r['file']}:{r['line']:>4} sectors {100*r['l2']/T['l2']:5.1f}%  x{r['l2']/max(r['l2i'],1):.2f} of ideal  {r['src']}")
