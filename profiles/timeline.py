"""clock64 timeline of the rot phases of chain 0 / CTA 0 / thread 0 (library built with -DPIMC_TIMELINE)."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
import numpy as np
pkg = ge.load_package()
geo = [int(x) for x in sys.argv[1:4]] + [0] * 3
cfg = pkg.configs.make_config(os.environ.get("TL_WORKLOAD", "C5"), worm=bool(os.environ.get("PROF_WORM")))
G = pkg.gpu.PimcGpu(cfg, nchains=int(os.environ.get("TL_CHAINS", "8")), ctas_per_chain=geo[0], threads_per_cta=geo[1], team=geo[2])
G.seed((12345,) * 6)
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 1024 + 5
G.steps(warm)
buf = (C.c_longlong * 4096)()
G.L.pimcgpu_timeline(buf, 4096)
G.steps(int(sys.argv[5]) if len(sys.argv) > 5 else 6)
n = G.L.pimcgpu_timeline(buf, 4096)
m = np.array(buf[:n], dtype=np.int64)
ids, t = m >> 48, m & 0xffffffffffff
names = {30: "batch geometry arrived", 31: "batch gathers consumed", 1: "rot sweep start", 2: "leader proposal done", 3: "after group sync", 4: "pair sums done (thread 0)", 5: "stage A done", 6: "after chain barrier",
         7: "leader decision done", 8: "after sync#3", 9: "decisions done (CTA 0)", 10: "after chain barrier",
         40: "worm: atom start", 41: "oc: gaussians", 42: "oc: set up by thread 0", 43: "oc: bridge filled", 44: "oc: written back", 45: "oc: potential sum", 46: "oc: decided",
         47: "ar: draws + gaussians", 48: "ar: set up by thread 0", 49: "ar: bridge filled", 50: "ar: written back", 51: "ar: potential sum", 52: "ar: decided",
         53: "swap: first table", 54: "swap: partner chosen", 55: "swap: gaussians", 56: "swap: plan", 57: "swap: bridge filled", 58: "swap: potential sum", 59: "swap: second table + decision", 61: "swap: done",
         20: "bisect: segment start", 21: "normals drawn", 22: "level pair sums done", 23: "level accept done", 24: "segment end", 25: "after chain barrier"}
prev = t[0]
for i in range(min(n, int(sys.argv[6]) if len(sys.argv) > 6 else 60)):
    print(f"{int(ids[i]):3d} {names.get(int(ids[i]), ''):28s} +{int(t[i]-prev):7d} cycles")
    prev = t[i]
G.close()
