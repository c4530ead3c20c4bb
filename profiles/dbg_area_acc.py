"""Accumulated space-fixed area tensor of the device (block accumulators after N measurements) against the oracle's sums over
the SAME trajectory (one chain, production cadence)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
from oracle import oracle_py as op
pkg = ge.load_package()
cfg = pkg.configs.make_config("C1", P=64, Q=16, temperature=1.0)
s = cfg.system
G = pkg.gpu.PimcGpu(cfg, nchains=1, chain_offset=1)
seed = (12345,) * 6
G.seed(seed)
O = op.Oracle(cfg)
O.sched_seed(seed, 1)
G.steps(200 * s.P); O.sched_run(0, 200 * s.P); t = 200 * s.P
G.accum_reset()
A = np.zeros(3); I = np.zeros(9); K = 0.0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for it in range(n):
    G.steps(16); G.measure()
    O.sched_run(t, 16); t += 16
    a3, i9 = O.area_estim3d(0)
    A += np.array(a3) ** 2; I += np.array(i9) / s.P; K += O.get_kin()
    O.sched_symmetry(s.reflect[0], s.reflect[1], s.reflect[2], 1 if s.rotsym else 0, max(1, s.rotsym))
acc, lay = G.accum_download()
a = acc[lay["area"]:lay["area"] + 36]
print("count", acc[0], "K gpu", acc[1] / acc[0], "oracle", K / n)
print("A^2 gpu   ", a[6 + 0], a[6 + 2], a[6 + 5])
print("A^2 oracle", A)
print("I   gpu   ", a[12 + 0], a[12 + 4], a[12 + 8])
print("I   oracle", I[0], I[4], I[8])
G.close()
