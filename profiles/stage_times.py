"""Per-stage wall times of the C5 step kernel: python profiles/stage_times.py [workload chains cpc threads team]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
w = sys.argv[1] if len(sys.argv) > 1 else "C5"
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 8
geo = [int(x) for x in sys.argv[3:6]] + [0] * 3
cfg = pkg.configs.make_config(w)
s = cfg.system
G = pkg.gpu.PimcGpu(cfg, nchains=chains, ctas_per_chain=geo[0], threads_per_cta=geo[1], team=geo[2])
G.seed((12345,) * 6)
print(G.geometry())
G.steps(s.P)          # warm-up pass
def timed(n):
    t = time.perf_counter(); G.steps(n); return (time.perf_counter() - t) * 1e6
nseg = [s.P // (1 << t.levels) for t in s.types]
t_first = timed(1)                  # time 0: molecular + bisection of every type + rot
rot_only = [k for k in range(1, s.P) if all(k % n for n in nseg)]
n = min(nseg) - 1
t_rot = timed(n) / n                # times 1..min(nseg)-1: rot sweep only
t_bis = timed(1)                    # time = min(nseg): bisection sweep (+ rot)
print(f"{w} chains={chains} geo={geo[:3]}: step@time0 {t_first:.0f} us | rot-only step {t_rot:.1f} us | step with bisection sweep {t_bis:.0f} us")
bu = s.bead_updates_per_pass()
est = t_rot * s.P + sum((t_bis - t_rot) * (1 << t.levels) for t in s.types) / len(s.types) + (t_first - t_bis)
print(f"   estimated pass {est/1e3:.1f} ms -> {bu['total']*chains/est*1e6/1e6:.1f} M bead-updates/s")
G.close()
