"""Free asymmetric top with device-generated tables: <E_rot> against the exact thermal energy (exploration script)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); gpu = pkg.gpu
T, Q = float(sys.argv[1]), int(sys.argv[2]); rtstep = float(sys.argv[3]); nch = int(sys.argv[4]); nblocks = int(sys.argv[5])
A, B, C = pkg.configs.ROT_CONSTANTS["H2O"]
maxj = gpu.asym_auto_maxj(T, Q, A, B, C)
r, e, q, info = gpu.gen_asymrho(T, Q, -1, 0, 180, A, B, C, maxj)
exact = info[7] / 0.6950356
cfg = pkg.configs.make_config("C4", P=64, Q=Q, big_tables=False, temperature=T)
cfg.system.types[0].numb = 1
cfg.system.types[0].rtstep = rtstep
cfg.coords, cfg.angles = pkg.configs.cluster_config(cfg.system, 3)
cfg.tables["rot3d"] = (r.reshape(-1), e.reshape(-1), q.reshape(-1))
G = gpu.PimcGpu(cfg, nchains=nch)
G.seed((4242,) * 6)
G.steps(400 * cfg.system.P)
rows = []
for b in range(nblocks):
    G.accum_reset()
    for k in range(250):
        G.steps(8, sync=False); G.measure()
    G.sync(); s = G.block_scalars(); rows.append(s.rot / s.count)
tot, acc = G.counters(); G.close()
rows = np.array(rows)
print("blocks", np.round(rows, 2))
print(f"T={T} Q={Q} rtstep={rtstep}: <E_rot> = {rows.mean():.4f} +- {rows.std(ddof=1)/np.sqrt(len(rows)):.4f} K, exact {exact:.4f} K (1.5 kT = {1.5*T}), maxj {maxj}, acceptance {acc[0][2]/max(tot[0][2],1):.3f}")
