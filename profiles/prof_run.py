"""Profiling driver: python profiles/prof_run.py <workload> <chains> <warm_steps> <nsteps> [cpc threads team]
launch 0 = warm-up (warm_steps), launch 1 = the launch to capture (nsteps, starting at time = warm_steps mod P)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
w, chains, warm, nsteps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
geo = [int(x) for x in sys.argv[5:8]] + [0] * 3
cfg = pkg.configs.make_config(w, worm=bool(os.environ.get("PROF_WORM")))      # PROF_WORM=1 keeps the deck's WORM line (C2, C3)
G = pkg.gpu.PimcGpu(cfg, nchains=chains, ctas_per_chain=geo[0], threads_per_cta=geo[1], team=geo[2])
G.seed((12345,) * 6)
G.steps(warm)
t = time.perf_counter(); G.steps(nsteps); dt = time.perf_counter() - t
print(f"{w} chains={chains} warm={warm} nsteps={nsteps} geo={geo[:3]} wall {dt*1e3:.2f} ms  -> {dt/nsteps*1e6:.1f} us/step")
G.close()
