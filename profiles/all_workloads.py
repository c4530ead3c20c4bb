"""Throughput of every BASELINE configuration on one GPU: python profiles/all_workloads.py [C1 C3 ...]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
ALL = (("C1", 148, 512), ("C2", 148, 512), ("C2w", 148, 512), ("C3", 148, 256), ("C3w", 148, 256), ("C4", 148, 64), ("C5", 8, 1024), ("C5", 15, 1024))
sel = set(sys.argv[1:])
for w, chains, nsteps in ALL:
    if sel and w not in sel:
        continue
    cfg = pkg.configs.make_config(w[:2], worm=w.endswith("w"))
    s = cfg.system
    G = pkg.gpu.PimcGpu(cfg, nchains=chains)
    G.seed((12345,) * 6)
    G.steps(max(8, nsteps // 8))
    G.accum_reset()
    t = time.perf_counter(); G.steps(nsteps); dt = time.perf_counter() - t
    bu = s.bead_updates_per_pass()
    tot, acc = G.counters()
    # bead-updates actually attempted in the window, from the move counters (SURVEY 8d weights)
    units = sum(tot[i, 0] * s.P + tot[i, 1] * ((1 << t_.levels) - 1) + tot[i, 2] for i, t_ in enumerate(s.types))
    if s.worm:
        wt, wa, cq = G.worm_counters()
        print(f"   worm acceptance open/close/advance/recede/swap {[round(float(a_ / max(t_, 1)), 3) for a_, t_ in zip(wa[[0, 1, 4, 5, 6]], wt[[0, 1, 4, 5, 6]])]}, {wt.sum() / dt / 1e6:.1f} M worm moves/s (not counted as bead-updates)")
    print(f"{w} chains={chains:3d} {G.geometry()} steps={nsteps} {dt*1e3:8.1f} ms -> {units/dt/1e6:8.1f} M bead-updates/s  acceptance {[round(float(a/max(t_,1)),3) for a, t_ in zip(acc.reshape(-1), tot.reshape(-1))]}")
    G.close()
