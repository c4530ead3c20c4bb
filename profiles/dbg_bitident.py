"""Dump the states of short runs of every small configuration (python profiles/dbg_bitident.py out.npz); two libraries
(PIMCGPU_LIB) that claim identical trajectories are compared with np.array_equal on the dumps."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
SMALL = dict(C5=dict(P=32, Q=8, nsolv=6), C4=dict(P=64, Q=32), C3=dict(P=32, Q=8), C1=dict(P=64, Q=16), C2=dict(P=32, Q=8, nsolv=3))
out = {}
for name, kw in SMALL.items():
    for worm in ((False, True) if name in ("C2", "C3") else (False,)):
        cfg = pkg.configs.make_config(name, worm=worm, **kw)
        G = pkg.gpu.PimcGpu(cfg, nchains=2)
        G.seed((21, 22, 23, 24, 25, 26))
        G.steps(3 * cfg.system.P + 7)
        for c in range(2):
            co, an, _ = G.download(c)
            out[f"{name}{'w' if worm else ''}_c{c}"] = co
            out[f"{name}{'w' if worm else ''}_a{c}"] = an
        G.close()
# full-size C1 and C2 as benchmarked (a few hundred steps)
for name in ("C1", "C2", "C3"):
    cfg = pkg.configs.make_config(name)
    G = pkg.gpu.PimcGpu(cfg, nchains=4)
    G.seed((31, 32, 33, 34, 35, 36))
    G.steps(200)
    co, an, _ = G.download(3)
    out[f"{name}_full_c"] = co; out[f"{name}_full_a"] = an
    G.close()
np.savez(sys.argv[1], **out)
print("saved", len(out))
