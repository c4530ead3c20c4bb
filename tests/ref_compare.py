"""Oracle port vs the reference's own C++ objects (oracle/_ref), one configuration per process.

Run as ``python tests/ref_compare.py <C1|C2|C3|C4|C5> [--golden OUT.npz]``.  Everything
compared here is CPU-only; the reference library keeps its state in globals so
tests/test_oracle_vs_ref.py runs this file once per configuration in a subprocess.
With --golden it also dumps inputs/outputs as a fixture the GPU box can check the
port against without /root/reference.
"""
import sys
import os
import json
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py as op  # noqa: E402

SMALL = dict(C5=dict(P=32, Q=8, nsolv=6), C4=dict(P=64, Q=32), C3=dict(P=32, Q=8), C1=dict(P=64, Q=16),
             C2=dict(P=32, Q=8, nsolv=3))


def rel(a, b):
    return abs(a - b) / max(1e-300, abs(b))


class _NoRef:
    """Stand-in when the reference objects are not available: every reference call returns None."""
    class _L:
        def __getattr__(self, k):
            return lambda *a, **kw: 0
    lib = _L()
    def rot_energy(self): return (0.0, 0.0, 0.0)
    def get_state(self): return None
    def push(self, *a): pass
    def queue_mode(self, *a): pass


def run(name, with_ref=True):
    """Returns (res, gold): comparisons against the reference (when with_ref) and the oracle's own outputs."""
    cfgs = op._configs()
    cfg = cfgs.make_config(name, **SMALL[name])
    s = cfg.system
    O = op.Oracle(cfg)
    R = op.Ref(cfg) if with_ref else _NoRef()
    rng = np.random.default_rng(12345)
    N, P, Q = s.N, s.P, s.Q
    res = {}
    gold = {}

    # --- spline set-up of the 1-D table (init_pot1D) and leaf interpolators -------------------
    if "pot1d" in cfg.tables:
        g = cfg.tables["pot1d"][0]
        rs = np.r_[rng.uniform(g[0] * 0.5, g[-1] * 1.3, 400), g[0], g[-1], g[5], g[17]]
        w = max(rel(O.spot1d(r)[0], R.lib.ref_SPot1D(r, 0)) for r in rs)
        res["SPot1D"] = w
        gold["spot1d_r"] = rs; gold["spot1d_v"] = np.array([O.spot1d(r)[0] for r in rs]); gold["spot1d_k"] = np.array([O.spot1d(r)[1] for r in rs])
    if "pot2d" in cfg.tables:
        rs = rng.uniform(1.5, 13.0, 400); cs = rng.uniform(-1.05, 1.05, 400)
        res["LPot2D"] = max(rel(O.lpot2d(r, c)[0], R.lib.ref_LPot2D(r, c, len(s.types) - 1)) for r, c in zip(rs, cs))
        gold["lpot2d_r"] = rs; gold["lpot2d_c"] = cs; gold["lpot2d_v"] = np.array([O.lpot2d(r, c)[0] for r, c in zip(rs, cs)])
    if "rotlin" in cfg.tables:
        gs = np.r_[rng.uniform(-1.02, 1.02, 400), -1.0, 1.0]
        it = len(s.types) - 1
        res["SRotDens"] = max(max(rel(O.srotdens(x, 0), R.lib.ref_SRotDens(x, it)), rel(O.srotdens(x, 1), R.lib.ref_SRotDensDeriv(x, it)),
                                  rel(O.srotdens(x, 2), R.lib.ref_SRotDensEsqrt(x, it))) for x in gs)
        gold["srot_g"] = gs; gold["srot_v"] = np.array([[O.srotdens(x, w) for w in range(3)] for x in gs])

    # --- per-bead potential sums -------------------------------------------------------------
    pe = np.array([[O.pot_energy_it(a, it) for it in range(P)] for a in range(N)])
    pr = np.array([[R.lib.ref_PotEnergy_it(a, it) for it in range(P)] for a in range(N)])
    res["PotEnergy_it"] = float(np.max(np.abs(pe - pr) / np.maximum(1e-300, np.abs(pr))))
    gold["pot_energy_it"] = pe
    pp = np.array([O.pot_energy_path(a) for a in range(N)])
    res["PotEnergy_path"] = max(rel(pp[a], R.lib.ref_PotEnergy_path(a)) for a in range(N))
    gold["pot_energy_path"] = pp
    mol = s.types[-1]
    if mol.molecule and Q:
        gm = N - mol.numb
        if mol.molecule == 2:
            e = np.array([1.1, 0.7, 2.3])
            v = [(O.pot_rot_e3d(gm, e, it), C_d(R.lib.ref_PotRotE3D, gm, e, it)) for it in range(P)]
        else:
            c = np.array([0.36, 0.48, 0.8])
            cc = np.zeros((3, N * P)); _, _, cs0 = O.get_state()
            v = [(O.pot_rot_energy(gm, cs0[:, gm * P + it // s.R], it), R.lib.ref_PotRotEnergy(gm, it)) for it in range(P)]
        res["PotRot"] = max(rel(a, b) for a, b in v)

    # --- estimators ----------------------------------------------------------------------------
    res["GetKin"] = rel(O.get_kin(), R.lib.ref_GetKinEnergy())
    res["GetPot"] = rel(O.get_pot(0), R.lib.ref_GetPotEnergy())
    O.reset_hist(); R.lib.ref_reset_block()
    res["GetPotDens"] = rel(O.get_pot(1), R.lib.ref_GetPotEnergy_Densities())
    gold["kin"] = O.get_kin(); gold["pot"] = O.get_pot(0)
    h = O.get_hist()
    if mol.molecule == 0 or any(t.molecule == 0 for t in s.types):
        g1 = np.zeros(300); R.lib.ref_get_gr1D(op._dp(g1)); res["gr1D"] = float(np.abs(g1 - h["gr1d"]).max())
    if mol.molecule == 1:
        g2 = np.zeros(300 * 50); R.lib.ref_get_gr2D(op._dp(g2)); res["gr2D"] = float(np.abs(g2 - h["gr2d"]).max())
        assert (not with_ref) or g2.sum() > 0
    if mol.molecule == 2 and len(s.types) > 1:
        g3 = np.zeros(300 * 50 * 100); R.lib.ref_get_gr3D(0, op._dp(g3)); res["gr3D"] = float(np.abs(g3 - h["gr3d_atoms"]).max())
        assert (not with_ref) or g3.sum() > 0
    if Q:
        if mol.molecule == 2:
            R.lib.ref_zero_relbins()
        a, b = O.get_rot_energy(), R.rot_energy()
        res["GetRot"] = max(rel(x, y) for x, y in zip(a, b))
        gold["rot"] = np.array(a)
        rc = np.zeros(Q); R.lib.ref_GetRCF(op._dp(rc))
        res["GetRCF"] = float(np.abs(rc - O.get_rcf()).max())
        gold["rcf"] = O.get_rcf()
        if mol.molecule == 2:
            h = O.get_hist()
            rt, rp, rch = np.zeros(50), np.zeros(100), np.zeros(100)
            R.lib.ref_get_relbins(op._dp(rt), op._dp(rp), op._dp(rch))
            res["relbins"] = float(max(np.abs(rt - h["relthe"]).max(), np.abs(rp - h["relphi"]).max(), np.abs(rch - h["relchi"]).max()))

    # --- area / exchange estimators (a18) on a permuted configuration -----------------------------
    bos = [t for t in s.types if t.stat == 1]
    if bos and mol.molecule:
        nb = bos[0].numb
        perm = np.roll(np.arange(nb, dtype=np.int32), 1) if nb > 1 else np.zeros(1, dtype=np.int32)
        if nb > 3:
            perm = np.arange(nb, dtype=np.int32); perm[[0, 1, 2]] = [1, 2, 0]       # one 3-cycle, the rest identity
        c0, a0, _ = O.get_state()
        O.set_state(c0, a0, perm)
        if with_ref:
            R.set_state(c0, a0, perm)
        pl = O.exchange_length(); gold["ploops"] = pl
        a3, i3 = O.area_estim3d(0); gold["area_sff"] = np.r_[a3, i3]
        if with_ref:
            rp = np.zeros(nb); R.lib.ref_exchange_length(op._dp(rp)); res["ploops"] = float(np.abs(rp - pl).max())
            ra, ri = np.zeros(6), np.zeros(9); R.lib.ref_area_estim3d(0, op._dp(ra), op._dp(ri))
            tri = np.array([a3[i] * a3[j] for i in range(3) for j in range(i + 1)])
            res["area_sff"] = float(max(np.abs(ra - tri).max() / max(1e-300, np.abs(tri).max()), np.abs(ri - i3).max() / np.abs(i3).max()))
        if mol.molecule == 2:
            a3, i3 = O.area_estim3d(1); gold["area_mff"] = np.r_[a3, i3]
            if with_ref:
                ra, ri = np.zeros(6), np.zeros(9); R.lib.ref_area_estim3d(1, op._dp(ra), op._dp(ri))
                tri = np.array([a3[i] * a3[j] for i in range(3) for j in range(i + 1)])
                res["area_mff"] = float(max(np.abs(ra - tri).max() / max(1e-300, np.abs(tri).max()), np.abs(ri - i3).max() / np.abs(i3).max()))
        else:
            a4 = O.area_estimators(); gold["area_lin"] = a4
            if with_ref:
                r4 = np.zeros(4); R.lib.ref_area_estimators(op._dp(r4))
                res["area_lin"] = float(np.max(np.abs(r4 - a4) / np.maximum(1e-300, np.abs(a4))))
        O.set_state(c0, a0, np.arange(nb, dtype=np.int32))
        if with_ref:
            R.set_state(c0, a0, np.arange(nb, dtype=np.int32))

    # --- moves with shared explicit uniforms -----------------------------------------------------
    R.queue_mode(True)
    nacc = mism = 0
    if Q:
        it_ = len(s.types) - 1
        for k in range(300):
            q = int(rng.integers(Q)); r = rng.random(4)
            if mol.molecule == 2:
                a0 = int(rng.integers(mol.numb))
                a = O.rot3d_step(q, a0, it_, r); b = R.lib.ref_MCRot3Dstep(q, a0, it_, *r)
            else:
                a = O.rotlin_step(q, it_, r[:3]); b = R.lib.ref_MCRotLinStep(q, it_, *r[:3])
            nacc += a; mism += (a != b)
        res["rot_steps_accepted"] = nacc
        res["rot_steps_decision_mismatch"] = mism
    nacc = mism = 0
    for k in range(200):
        typ = int(rng.integers(len(s.types))); t = s.types[typ]; time = int(rng.integers(P))
        nm = (1 << t.levels) - 1
        ug = rng.random(t.numb * nm * 6); ua = rng.random(t.numb * t.levels)
        R.lib.ref_rng_clear()
        R.push(8, ug[0::2]); R.push(9, ug[1::2]); R.push(3, ua)
        ig = ia = 0
        for a_ in range(t.numb):
            acc, cg, ca = O.bisection_move(typ, a_, time, ug[ig:], ua[ia:], 0)
            ig += cg; ia += ca; nacc += acc
        R.lib.ref_MCBisectionMove(typ, time)
        mism += (R.lib.ref_rng_pending(3) != len(ua) - ia) + (R.lib.ref_rng_pending(8) != (len(ug) - ig) // 2)
    res["bisection_accepted"] = nacc
    res["bisection_stream_mismatch"] = mism
    nacc = 0
    for k in range(10):
        typ = int(rng.integers(len(s.types))); t = s.types[typ]
        u = rng.random(t.numb * 3); ua = np.full(t.numb, rng.random())
        R.lib.ref_rng_clear(); R.push(1, u); R.push(2, ua)
        for a_ in range(t.numb):
            nacc += O.molecular_move(typ, a_, u[3 * a_:3 * a_ + 3], ua[0])
        R.lib.ref_MCMolecularMove(typ)
    res["molecular_accepted"] = nacc
    # --- symmetry operations (a19) on the moved state ------------------------------------------------
    if Q and mol.molecule:
        if mol.molecule == 2:
            for plane in (0, 1, 2):
                O.reflect(plane); R.lib.ref_reflect(plane)
        nfold = 2 if mol.molecule == 2 else 1
        R.lib.ref_set_nfold(nfold)
        for u in (0.3, 0.95):
            O.rotsym(u, nfold)
            R.lib.ref_rng_clear(); R.push(1, [u]); R.lib.ref_rotsym()
    co, ao, cso = O.get_state()
    cr, ar, csr = R.get_state() if with_ref else (co, ao, cso)
    res["state_cosine_maxdiff"] = float(np.abs(cso - csr).max())
    res["state_coords_maxdiff"] = float(np.abs(co - cr).max())
    res["state_angles_maxdiff"] = float(np.abs(ao - ar).max())
    tot, acc = O.counters()
    rt_, ra_ = np.zeros(6), np.zeros(6)
    R.lib.ref_counters(op._dp(rt_), op._dp(ra_))
    nt = len(s.types)   # rotational counters live in the harness' locals, compare MCMOLEC/MCMULTI only
    res["counters_maxdiff"] = float(max(np.abs(tot[:nt, :2] - rt_.reshape(2, 3)[:nt, :2]).max(),
                                        np.abs(acc[:nt, :2] - ra_.reshape(2, 3)[:nt, :2]).max()))
    res["final_kin"] = rel(O.get_kin(), R.lib.ref_GetKinEnergy())
    res["final_pot"] = rel(O.get_pot(0), R.lib.ref_GetPotEnergy())
    gold["final_coords"] = co; gold["final_angles"] = ao
    gold["final_kin"] = O.get_kin(); gold["final_pot"] = O.get_pot(0)
    return res, gold


def main():
    name = sys.argv[1]
    golden = sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--golden" else None
    res, gold = run(name, with_ref=True)
    print("RESULT " + json.dumps(res))
    if golden:
        np.savez_compressed(golden, **gold)


def C_d(fn, atom, e, it):
    import ctypes as C
    fn.restype = C.c_double
    e = np.ascontiguousarray(e, dtype=np.float64)
    return fn(C.c_int(atom), op._dp(e), C.c_int(it))


if __name__ == "__main__":
    main()
