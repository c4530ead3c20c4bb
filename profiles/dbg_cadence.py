"""Device trajectory against the oracle replay under the CADENCE of a production run: many short launches with a
measurement (estimators + symmetry moves) after each.  python profiles/dbg_cadence.py <config> <iters> <skip> [kw json]"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
from oracle import oracle_py as op
pkg = ge.load_package()
name, iters, skip = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
kw = json.loads(sys.argv[4]) if len(sys.argv) > 4 else {}
cfg = pkg.configs.make_config(name, **kw)
s = cfg.system
G = pkg.gpu.PimcGpu(cfg, nchains=3, chain_offset=0)
seed = (12345,) * 6
G.seed(seed)
O = op.Oracle(cfg)
O.sched_seed(seed, 1)
t = 0
for it in range(iters):
    G.steps(skip); G.measure()
    O.sched_run(t, skip); t += skip
    if any(s.reflect) or s.rotsym:
        O.sched_symmetry(s.reflect[0], s.reflect[1], s.reflect[2], 1 if s.rotsym else 0, max(1, s.rotsym))
    if it % max(1, iters // 20) == 0 or it == iters - 1:
        cg, ag, _ = G.download(1)
        co, ao, _ = O.get_state()
        rows = slice((s.N - 1) * s.P, (s.N - 1) * s.P + s.Q)
        print(f"iter {it:5d} step {t:7d}: max|dcoords| {np.abs(cg - co).max():.3e}  max|dangles| {np.abs(ag[:, rows] - ao[:, rows]).max():.3e}", flush=True)
G.close()
