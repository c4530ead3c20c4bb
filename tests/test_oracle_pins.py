"""CPU pins of the ORACLE's restatement of the Fortran leaves (the Fortran itself cannot be compiled in the image):
every leaf is checked against something that does not share code with oracle/fortran_shim.cpp --
 * caleng_  : a first-principles numpy TIP4P from the model's definition;
 * rotpro   : numpy forward-difference interpolation on the reference's own table plane nmv_prop/rho.den010_*;
 * vcord_   : numpy projection of the particle onto the matpre body axes, rigid-rotation and translation invariance;
 * vspher_  : the DATA table re-extracted from vspher.f when the reference tree is present; REAL*4 literal semantics;
 * deleul   : numpy R1^T R2 and its ZYZ angles.
These bound what "parity unpinned" leaves open for the Fortran leaves."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rotmat(e):
    """matpre (rotden.f:136-163): R = Rz(phi) Ry(theta) Rz(chi), written out independently in numpy"""
    def rz(a):
        return np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])

    def ry(a):
        return np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    return rz(e[0]) @ ry(e[1]) @ rz(e[2])


def _unit(rng, n):
    u = rng.standard_normal((n, 3))
    return u / np.linalg.norm(u, axis=1)[:, None]


def _euler(rng, n):
    return np.c_[rng.uniform(0, 2 * np.pi, n), np.arccos(rng.uniform(-1, 1, n)), rng.uniform(0, 2 * np.pi, n)]


@pytest.fixture(scope="module")
def oc1(pkg):
    from oracle import oracle_py as op
    cfg = pkg.configs.make_config("C1", P=64, Q=16)
    return cfg, op.Oracle(cfg)


def test_caleng_against_first_principles_tip4p(pkg):
    from oracle import oracle_py as op
    cfg = pkg.configs.make_config("C4", P=64, Q=32, big_tables=False)
    O = op.Oracle(cfg)
    rng = np.random.default_rng(21)
    n = 2000
    c1 = rng.uniform(-1, 1, (n, 3)); c2 = c1 + rng.uniform(2.4, 8.0, (n, 1)) * _unit(rng, n)
    e1, e2 = _euler(rng, n), _euler(rng, n)
    body = {"O": np.array([0, 0, 0.06562]), "H1": np.array([0.7557, 0, -0.5223]), "H2": np.array([-0.7557, 0, -0.5223]), "M": np.array([0, 0, -0.08438])}
    q = {"H1": 0.52, "H2": 0.52, "M": -1.04}
    coul = 3.1577465e5 * 0.52917721092          # hartree (K) x bohr (Angstrom): e^2 / (4 pi eps0)
    worst = 0.0
    for i in range(n):
        Ra, Rb = _rotmat(e1[i]), _rotmat(e2[i])
        sa = {k: c1[i] + Ra @ v for k, v in body.items()}
        sb = {k: c2[i] + Rb @ v for k, v in body.items()}
        roo = np.linalg.norm(sa["O"] - sb["O"])
        e = (6.0e5 / roo ** 12 - 610.0 / roo ** 6) * 503.218978939
        for ka, qa in q.items():
            for kb, qb in q.items():
                e += qa * qb * coul / np.linalg.norm(sa[ka] - sb[kb])
        worst = max(worst, abs(O.caleng(c1[i], c2[i], e1[i], e2[i]) - e) / max(abs(e), 1.0))
    assert worst < 1e-10, worst          # nine Coulomb terms of ~1e4 K and both signs: 1e-16 x 1e5 K against max(|E|, 1 K)


def test_rotpro_on_the_references_table_plane(pkg, oc1):
    from oracle import oracle_py as op
    z = np.load(os.path.join(ROOT, "tests", "golden", "tablegen", "ref_asymrho_den010.npz"))
    cfg = pkg.configs.make_config("C1", P=64, Q=16)
    tabs = [t.reshape(181, 361, 361).copy() for t in cfg.tables["rot3d"]]
    for t, k in zip(tabs, ("rho", "eng", "esq")):
        t[10] = z[k]; t[11] = z[k] * 1.03
    cfg.tables["rot3d"] = tuple(np.ascontiguousarray(t.reshape(-1)) for t in tabs)
    O = op.Oracle(cfg)
    rng = np.random.default_rng(5)
    n = 5000
    deg = np.c_[rng.uniform(0, 360, n), rng.uniform(10, 11, n), rng.uniform(0, 360, n)]
    o = np.array([O.rotpro(d) for d in deg])
    ip, it, ic = deg[:, 0].astype(int), deg[:, 1].astype(int), deg[:, 2].astype(int)
    assert np.array_equal(o[:, 3].astype(int), (it * 361 + ip) * 361 + ic)        # theta outer, phi, chi inner (rotpro_sub.f:26)
    for col, t in ((0, tabs[0]), (1, tabs[1]), (2, tabs[2])):
        f0 = t[it, ip, ic]
        ind = f0 + (t[it, ip, ic + 1] - f0) * (deg[:, 2] - ic) + (t[it, ip + 1, ic] - f0) * (deg[:, 0] - ip) + (t[it + 1, ip, ic] - f0) * (deg[:, 1] - it)
        assert np.max(np.abs(o[:, col] - ind) / np.maximum(np.abs(ind), 1e-12 * np.abs(t[10]).max())) < 1e-13
    # last grid lines: zero difference along the saturated axis
    r = O.rotpro([360.0, 10.5, 17.25])
    f0 = tabs[0][10, 360, 17]
    assert abs(r[0] - (f0 + (tabs[0][10, 360, 18] - f0) * 0.25 + (tabs[0][11, 360, 17] - f0) * 0.5)) < 1e-13 * abs(f0)


def test_deleul_is_the_zyz_decomposition_of_r1t_r2(oc1):
    _, O = oc1
    rng = np.random.default_rng(8)
    for e1, e2 in zip(_euler(rng, 500), _euler(rng, 500)):
        rel = O.deleul(e1, e2)
        M = _rotmat(e1).T @ _rotmat(e2)
        assert np.max(np.abs(_rotmat(rel) - M)) < 1e-12
        assert 0 <= rel[0] < 2 * np.pi + 1e-12 and 0 <= rel[1] <= np.pi and 0 <= rel[2] < 2 * np.pi + 1e-12


def test_vcord_projection_and_invariances(oc1):
    cfg, O = oc1
    rng = np.random.default_rng(33)
    n = 1500
    eul = _euler(rng, n)
    rcom = rng.uniform(-1, 1, (n, 3))
    d = rng.uniform(2.5, 8.0, (n, 1)) * _unit(rng, n)
    S = _rotmat(np.array([0.7, 1.1, 2.3]))
    for i in range(n):
        v0, rtc0, i0 = O.vcord(eul[i], rcom[i], rcom[i] + d[i])
        R = _rotmat(eul[i])
        body = R.T @ d[i]
        r = np.linalg.norm(d[i])
        assert abs(rtc0[0] - r) < 1e-12
        assert abs(rtc0[1] - np.arccos(np.clip(body[2] / r, -1, 1))) < 1e-9
        dch = abs(rtc0[2] - np.mod(np.arctan2(body[1], body[0]), 2 * np.pi))
        assert min(dch, 2 * np.pi - dch) < 1e-9
        # table index from (r, theta, chi) by the documented layout: r outer (bohr), theta, chi inner, chi folded about pi (181 grid)
        chi = rtc0[2] if rtc0[2] <= np.pi else 2 * np.pi - rtc0[2]
        rb = min(max(r / 0.529177249, 4.0), 20.0)
        ir, ith, ich = int((rb - 4.0) / (16.0 / 500)), min(int(np.degrees(rtc0[1])), 180), min(int(np.degrees(chi)), 180)
        if abs(np.degrees(rtc0[1]) - round(np.degrees(rtc0[1]))) > 1e-9 and abs(np.degrees(chi) - round(np.degrees(chi))) > 1e-9 \
                and abs((rb - 4.0) / 0.032 - round((rb - 4.0) / 0.032)) > 1e-9:
            assert i0 == (ir * 181 + ith) * 181 + ich
        # translation of the whole system
        sh = rng.uniform(-30, 30, 3)
        v1, rtc1, i1 = O.vcord(eul[i], rcom[i] + sh, rcom[i] + sh + d[i])
        assert np.max(np.abs(rtc1 - rtc0)) < 1e-12
        # rigid rotation of the whole system
        R2 = S @ R
        e2 = np.array([np.mod(np.arctan2(R2[1, 2], R2[0, 2]), 2 * np.pi), np.arccos(np.clip(R2[2, 2], -1, 1)), np.mod(np.arctan2(R2[2, 1], -R2[2, 0]), 2 * np.pi)])
        v2, rtc2, i2 = O.vcord(e2, rcom[i], rcom[i] + S @ d[i])
        st = max(np.sin(rtc0[1]), 1e-6)
        dd = np.abs(rtc2 - rtc0); dd[2] = min(dd[2], 2 * np.pi - dd[2])
        assert dd[0] < 1e-12 and dd[1] * st < 1e-12 and dd[2] * st < 1e-11
        if i2 == i0:
            assert abs(v2 - v0) <= 1e-8 * max(abs(v0), 1e-6)


def test_vspher_table_and_literal_semantics(pkg):
    from oracle import oracle_py as op
    t = pkg.configs.load_vspher_table()
    g = np.loadtxt(os.path.join(ROOT, "tests", "golden", "vspher_table.txt"))
    assert t.shape == (501,) and np.array_equal(t, g)
    assert np.array_equal(t, t.astype(np.float32).astype(np.float64))          # REAL*4 literals widened to double
    src = "/root/reference/vspher.f"
    if os.path.exists(src):
        import re
        lits = re.findall(r"^\s+\+\s*([-+]?\d+\.\d+E[-+]\d+)\s*[,/]", open(src).read(), flags=re.M)
        assert len(lits) == 501
        assert np.array_equal(np.array([float(x) for x in lits], dtype=np.float32).astype(np.float64), t)
    cfg = pkg.configs.make_config("SPH", P=32, nsolv=2)
    O = op.Oracle(cfg)
    a2b = float(np.float32(0.5291772))
    # on grid points the interpolation returns the table entry; outside the range the value is clamped and so is r
    for i in (0, 1, 250, 499, 500):
        rb = 3.0 + 0.046 * i
        v, rc = O.vspher(rb * a2b)
        assert abs(rc - rb) < 1e-12 and abs(v - t[i]) <= 1e-9 * abs(t[i])
    assert O.vspher(0.1) == (t[0], 3.0) and O.vspher(40.0) == (t[500], 26.0)
    v, rc = O.vspher((3.0 + 0.046 * 10.5) * a2b)
    assert abs(v - 0.5 * (t[10] + t[11])) < 1e-9 * abs(t[10])


def test_new_selectors_are_exported(pkg):
    import ctypes
    L = ctypes.CDLL(pkg.gpu.LIB)
    for name in ("pimcgpu_eval_rotpro", "pimcgpu_eval_vcalc", "pimcgpu_eval_deleul", "pimcgpu_eval_vcord_grid", "pimcgpu_eval_vspher", "pimcgpu_eval_libm"):
        assert hasattr(L, name)
