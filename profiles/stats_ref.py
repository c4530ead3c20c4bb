"""Converged-estimator fixtures from the REFERENCE's own sampling (oracle/_ref: its unmodified move and estimator objects).

    python profiles/stats_ref.py [case ...]        (run in the build container, where /root/reference exists)

writes tests/golden/stats/<case>_ref.json: one row per block with the columns of observables() below -- <K>, <V>, <E_rot>,
the orientational correlation <n(0).n(t)> at t = 1, Q/4, Q/2 (GetRCF, mc_estim.cc:1099-1139, per time origin), and the
superfluid fractions: linear dopant _area2*norm/_inert perp/parallel (.sup columns 2-3, mc_estim.cc:2626-2627); top
4m^2<A_iA_i>/(beta hbar^2 I_ii) in the space-fixed and dopant-fixed frames (.sffs3d/.mffs3d, mc_estim.cc:2696-2718).
Each case runs in its own process (the reference keeps its state in globals)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# name -> (config, make_config kwargs, blocks, steps per block, steps between measurements, equilibration passes)
CASES = {
    "top_He_C1_P64_Q16_1K": ("C1", dict(P=64, Q=16, temperature=1.0), 64, 25600, 16, 3000),
    "tip4p_C4_P64_Q32": ("C4", dict(P=64, Q=32), 64, 25600, 16, 3000),
    "lin_C5_P64_Q16_6H2_2K": ("C5", dict(P=64, Q=16, nsolv=6, temperature=2.0), 128, 25600, 16, 3000),
}


def observables(s, n, k, v, e, rcf, lin6, sff15, mff15, lam_b, mass_b):
    """one block row from raw block sums (n = measurements in the block); shared by the GPU tests"""
    Q = max(1, s.Q)
    beta = 1.0 / s.temperature
    row = [k / n, v / n, e / n]
    for t in (1, Q // 4, Q // 2):
        row.append(rcf[t] / (n * Q) if s.Q else 0.0)
    mol = [t for t in s.types if t.molecule]
    bose = any(t.stat == 1 for t in s.types)
    if bose and mol and mol[0].molecule == 1:            # SaveAreaEstimators: _area2*norm/_inert, norm = 2/(beta lambda)
        norm = 2.0 / (beta * lam_b)
        row += [lin6[2] * norm / lin6[4], lin6[3] * norm / lin6[5]]
    else:
        row += [0.0, 0.0]
    if bose:                                              # SaveAreaEstim3D: norm = 2 m/(beta lambda); diagonal ids 0, 2, 5 / 0, 4, 8
        norm = 2.0 * mass_b / (beta * lam_b)
        for fr in (sff15, mff15):
            for ia, ii in ((0, 0), (2, 4), (5, 8)):
                row.append(fr[ia] * norm / fr[6 + ii] if fr[6 + ii] != 0.0 else 0.0)
    else:
        row += [0.0] * 6
    return row


def blocked_sem(x):
    """standard error of the mean of a correlated series by the blocking method: blocks are merged pairwise while at
    least eight remain and the LARGEST estimate is kept (slow modes make single-level estimates too small)"""
    y = np.asarray(x, dtype=float)
    best = 0.0
    while len(y) >= 8:
        best = max(best, y.std(ddof=1) / np.sqrt(len(y)))
        y = y[:len(y) // 2 * 2].reshape(-1, 2).mean(axis=1)
    return best


COLS = ["K", "V", "E_rot", "rcf(1)", "rcf(Q/4)", "rcf(Q/2)", "fs_perp(.sup)", "fs_par(.sup)", "fs_xx(sff)", "fs_yy(sff)", "fs_zz(sff)",
        "fs_xx(mff)", "fs_yy(mff)", "fs_zz(mff)"]


def run_case(case):
    from oracle import oracle_py as op
    name, kw, nblocks, per_block, skip, eq = CASES[case]
    cfg = op._configs().make_config(name, **kw)
    s = cfg.system
    R = op.Ref(cfg)
    L = R.lib
    L.ref_lambda.argtypes = [__import__("ctypes").c_int]
    bt = [i for i, t in enumerate(s.types) if t.stat == 1]
    lam_b = L.ref_lambda(bt[0]) if bt else 1.0
    mass_b = s.types[bt[0]].mass if bt else 1.0
    L.ref_run_steps(0, eq * s.P)
    t = eq * s.P
    out7 = np.zeros(7); rcf = np.zeros(max(1, s.Q)); lin6 = np.zeros(6); sff = np.zeros(15); mff = np.zeros(15)
    rows = []
    for b in range(nblocks):
        L.ref_reset_block()
        n = 0
        for _ in range(per_block // skip):
            L.ref_run_steps(t, skip); t += skip
            L.ref_MCGetAverage(op._dp(out7)); n += 1
        L.ref_get_block_acc(op._dp(rcf), op._dp(lin6), op._dp(sff), op._dp(mff))
        rows.append(observables(s, n, out7[0], out7[1], out7[2], rcf, lin6, sff, mff, lam_b, mass_b))
    print("ROWS " + json.dumps({"case": case, "config": name, "kw": kw, "per_block": per_block, "skip": skip, "columns": COLS,
                                "lambda_bose": lam_b, "mass_bose": mass_b, "rows": rows}))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        run_case(sys.argv[2])
        sys.exit(0)
    for case in (sys.argv[1:] or list(CASES)):
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", case], capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("ROWS ")]
        if not line:
            print(out.stdout[-2000:], out.stderr[-2000:]); raise SystemExit(f"{case}: no output")
        d = json.loads(line[-1][5:])
        path = os.path.join(ROOT, "tests", "golden", "stats", case + "_ref.json")
        json.dump(d, open(path, "w"))
        r = np.array(d["rows"])
        print(case, "->", path)
        for i, c in enumerate(COLS):
            print(f"   {c:16s} {r[:, i].mean(): .6g} +- {blocked_sem(r[:, i]):.3g}")
