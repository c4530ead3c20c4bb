import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package()
import test_gpu_worm as T
from oracle import oracle_py as op
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = T.make(pkg, name); s = cfg.system
s.worm = (s.worm[0], 0.003, s.worm[2])
nb = s.types[0].numb
G = pkg.gpu.PimcGpu(cfg, nchains=2); O = op.Oracle(cfg)
print(G.geometry())
seed = (77, 78, 79, 80, 81, 82)
G.seed(seed); O.sched_seed(seed, 1)
rows = T.rotor_rows(s)
for t in range(3 * s.P + 2):
    G.steps(1); O.sched_run(t, 1)
    cg, ag, _ = G.download(1); co, ao, _ = O.get_state()
    dc = np.abs(cg - co).max(); da = np.abs(ag[:, rows] - ao[:, rows]).max()
    ws, wo = G.worm_state(1), O.worm_get()
    pg, po = G.download_perm(1), O.get_perm(nb)[0]
    bad = dc > 1e-8 or da > 1e-8 or ws != wo or not np.array_equal(pg, po)
    if bad or t % 16 == 0:
        print(t, "time", t % s.P, "dc %.2e da %.2e" % (dc, da), ws, wo, pg.tolist(), po.tolist(), O.worm_counters()[1].tolist())
    if bad:
        d = np.abs(cg - co).reshape(3, s.N, s.P).max(axis=0)
        print("coords diff per atom:", d.max(axis=1)); print("slices with diff:", np.where(d.max(axis=0) > 1e-8)[0])
        da_ = np.abs(ag - ao)[:, rows].max(axis=0); print("angle diff per rot slice", da_)
        print("counters G", G.counters()[0].tolist(), "O", O.counters()[0].tolist())
        break
G.close()
