"""Where the end-to-end step of bench.py goes (C ABI with host buffers): wall time per phase, synchronised after each."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
w = sys.argv[1] if len(sys.argv) > 1 else "C5"
chains = int(sys.argv[2]) if len(sys.argv) > 2 else (8 if w == "C5" else 148)
cfg = pkg.configs.make_config(w)
s = cfg.system
G = pkg.gpu.PimcGpu(cfg, nchains=chains)
G.seed((12345,) * 6)
n = s.N * s.P
pin_c = torch.empty((chains, 3, n), dtype=torch.float64, pin_memory=True)
pin_a = torch.empty((chains, 3, n), dtype=torch.float64, pin_memory=True)
host_c, host_a = pin_c.numpy(), pin_a.numpy()
G.download_rows_into(host_c, host_a)
G.steps(s.P)
t = {k: 0.0 for k in ("upload", "steps", "measure", "accum", "download")}
reps = 5
for _ in range(reps):
    t0 = time.perf_counter()
    G.upload_all(host_c, host_a, cfg.perm)
    G.sync(); t1 = time.perf_counter()
    G.accum_reset(); G.steps(s.P, sync=False); G.sync(); t2 = time.perf_counter()
    G.measure(); G.sync(); t3 = time.perf_counter()
    acc, _ = G.accum_download(); t4 = time.perf_counter()
    G.download_rows_into(host_c, host_a)
    t5 = time.perf_counter()
    for k, v in zip(t, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
        t[k] += v
print(w, chains, "chains; ms per step:", {k: round(1e3 * v / reps, 3) for k, v in t.items()}, "total", round(1e3 * sum(t.values()) / reps, 2))
G.close()
