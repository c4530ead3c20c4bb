"""GPU-side worm statistics: python profiles/worm_stats_gpu.py <nblocks> <nchains> <out.json>"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package()
nblocks, nchains = int(sys.argv[1]), int(sys.argv[2])
cfg = pkg.configs.make_config("C2", P=64, Q=16, nsolv=5, temperature=1.0)
cfg.system.worm = ("He4", 0.13, 8); cfg.system.reflect, cfg.system.rotsym = (0, 0, 0), 0
nb = cfg.system.types[0].numb
G = pkg.gpu.PimcGpu(cfg, nchains=nchains)
G.seed((12345,) * 6)
G.steps(200 * cfg.system.P)
rows = []
for b in range(nblocks):
    G.accum_reset()
    for k in range(3200 // 16):
        G.steps(16, sync=False); G.measure()
    G.sync()
    acc, lay = G.accum_download()
    n = acc[0]; pl = acc[lay["ploops"]:lay["ploops"] + nb]; a6 = acc[lay["area"] + 6:lay["area"] + 12]
    rows.append([acc[1] / n, acc[2] / n, float(sum((l + 1) * pl[l] for l in range(1, nb)) / (n * nb)), (a6[0] + a6[2] + a6[5]) / n, n])
wt, wa, _ = G.worm_counters(); t, a = G.counters()
print("worm acceptance", (wa / np.maximum(wt, 1)).round(3).tolist(), "moves", (a / np.maximum(t, 1)).round(3).tolist())
json.dump(rows, open(sys.argv[3], "w"))
g = np.array(rows)
for i, nm in enumerate(("K", "V", "exch", "A.A", "n")):
    print(nm, g[:, i].mean(), g[:, i].std(ddof=1) / np.sqrt(len(g)))
