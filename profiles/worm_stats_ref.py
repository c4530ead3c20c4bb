"""Reference-side worm statistics (CPU): python profiles/worm_stats_ref.py <nblocks> <out.json>"""
import sys, re, subprocess, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "tests", "test_gpu_statistics.py")).read()
code = re.search(r"WORM_REF_SCRIPT = r'''(.*?)'''", src, re.S).group(1) % dict(root=ROOT, name="C2", kw=dict(P=64, Q=16, nsolv=5, temperature=1.0),
                                                                               worm=("He4", 0.13, 8), nblocks=int(sys.argv[1]), per_block=12800, skip=16)
out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
line = [l for l in out.stdout.splitlines() if l.startswith("ROWS ")][-1]
open(sys.argv[2], "w").write(line[5:])
