import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge
g = ge.load_package().gpu
t = time.time()
r, e, q, info = g.gen_asymrho(0.37, 16384, -1, 0, 0, 0.6666525, 0.2306476, 0.1769383, 876)
print(876, 'rho(identity)*8pi^2', r[0, 0, 0] * 8 * np.pi ** 2, 'Z(tau)', info[13], 'E(identity)', e[0, 0, 0], 'E(tau)', info[14], 'finite', np.isfinite(r).all(), round(time.time() - t, 2), 's', g.gen_timing())
