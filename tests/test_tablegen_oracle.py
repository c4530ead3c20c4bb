"""CPU tests pinning the table-generator oracle (oracle/tablegen_oracle.cpp) on the reference's OWN golden outputs
(fixtures tests/golden/tablegen/ref_*.npz = nmv_prop/rho.den010_*, nmv_prop/log, symtop_prop/rho.den0{00,10}_*,
examples/*/N2O_T0.5t128.rot, CO2_T100t4.rot; made by tests/golden/tablegen/make_fixtures.py).

Bars: linden.f -- BYTE-identical files (the arithmetic is double precision and order-preserving);
asymrho.f / symrho.f -- the tables are printed with 8 significant digits (E15.8) and carry their own rounding
noise of ~1e-15 * max|rho| from the alternating sums, so |delta| <= half a unit of the last printed digit + noise where rho is above the noise."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "tablegen")


def ulp8(g):
    """one unit in the last of the 8 digits E15.8 prints"""
    g = np.abs(np.asarray(g, dtype=float))
    return 10.0 ** (np.floor(np.log10(np.maximum(g, 1e-300))) + 1 - 8)


@pytest.fixture(scope="module")
def tg():
    from oracle import tablegen_py
    return tablegen_py


def test_fortran_e15_8_edit_descriptor(tg):
    assert tg.fmt_e15_8(6.2732329) == " 0.62732329E+01"
    assert tg.fmt_e15_8(-0.50630561) == "-0.50630561E+00"
    assert tg.fmt_e15_8(0.0) == " 0.00000000E+00"
    assert tg.fmt_e15_8(-1437.6686) == "-0.14376686E+04"
    assert tg.fmt_e15_8(9.99999999e-5) == " 0.10000000E-03"          # rounding carries into the exponent
    assert tg.fmt_e15_8(1.5e-120) == " 0.15000000-119"               # three-digit exponents drop the E
    assert tg.fmt_e15_8(-1.0, True) == "-1.00000000E+00"
    assert tg.fmt_e15_8(8.47475547, True) == " 8.47475547E+00"


@pytest.mark.parametrize("name", ["N2O", "CO2"])
def test_linden_oracle_is_byte_identical_to_reference_rot_file(tg, name):
    f = np.load(os.path.join(GOLD, f"ref_linden_{name}.npz"))
    T, ns, B, npt, io = f["args"]
    out, info = tg.linden(float(T), int(ns), float(B), int(npt), int(io))
    assert tg.rot_lines(out) == tg.rot_lines(f["table"])             # every line of the reference's file
    assert np.array_equal(out, out) and out.shape == (int(npt), 4)
    assert info[1] > 10


def test_asymrho_oracle_partition_functions_match_reference_log(tg):
    f = np.load(os.path.join(GOLD, "ref_asymrho_den010.npz"))
    T, ns, io, _, _, A, B, C, maxj = f["args"]
    o = tg.AsymRho(float(T), int(ns), int(io), float(A), float(B), float(C), int(maxj))
    i = o.info
    # log prints F12.6: Z, E (cm-1), E (K), Cv for even k / odd k / classical (asymrho.f:354-368)
    mine = np.array([[i[0], i[1], i[1] / 0.6950356, i[2]], [i[3], i[4], i[4] / 0.6950356, i[5]], [i[6], i[7], i[7] / 0.6950356, i[8]]])
    assert np.all(np.abs(mine - f["at_beta"]) <= 0.5000001e-6)
    mine_tau = np.array([[i[9], i[10], i[10] / 0.6950356], [i[11], i[12], i[12] / 0.6950356], [i[13], i[14], i[14] / 0.6950356]])
    assert np.all(np.abs(mine_tau - f["at_tau"]) <= 0.5000001e-6 * np.maximum(1.0, np.abs(f["at_tau"])) * 10)
    o.close()


def test_asymrho_oracle_matches_reference_plane_samples(tg):
    """asymrho.f at theta = 10 deg, maxj = 66: ~1e7 cosine terms per grid point on the CPU, so a sample of the plane."""
    f = np.load(os.path.join(GOLD, "ref_asymrho_den010.npz"))
    T, ns, io, ith, _, A, B, C, maxj = f["args"]
    o = tg.AsymRho(float(T), int(ns), int(io), float(A), float(B), float(C), int(maxj))
    rmax = np.abs(f["rho"]).max()
    pts = [(0, 0), (1, 0), (3, 2), (7, 7), (12, 5), (20, 10), (359, 1), (355, 4), (340, 15), (30, 0)]
    for (ip, ic) in pts:
        v = o.point(int(ith), ip, ic)
        g = np.array([f["rho"][ip, ic], f["eng"][ip, ic], f["esq"][ip, ic]])
        assert abs(g[0]) > 1e-7 * rmax, "sample point inside the reference's own rounding noise"
        assert np.all(np.abs(v - g) <= 0.51 * ulp8(g) + 1e-13 * rmax * np.array([1, 0, 0])), (ip, ic, v, g)
    o.close()


def test_asymrho_symmetry_fill_is_an_involution_free_gather(tg):
    """the four sequential passes (asymrho.f:665-709) only copy from the directly computed region chi <= maxchi(phi)"""
    idx = np.arange(361 * 361, dtype=np.float64).reshape(361, 361).copy()
    tg.lib().tg_asym_symfill(idx.ctypes.data_as(tg.c_dp))
    src = idx.astype(np.int64)
    si, sl = src // 361, src % 361
    direct = np.array([[ichi <= tg.maxchi(iphi) for ichi in range(361)] for iphi in range(361)])
    assert direct[si, sl].all()
    assert np.array_equal(src, src.T)                       # rho(phi,chi) = rho(chi,phi) after the last pass


@pytest.mark.parametrize("ith", [0, 10])
def test_symrho_oracle_matches_reference_planes(tg, ith):
    f = np.load(os.path.join(GOLD, f"ref_symrho_den{ith:03d}.npz"))
    T, ns, kmod, _, _, Bz, Bxy, maxj = f["args"]
    r, e, q, info = tg.symrho_plane(float(T), int(ns), int(kmod), ith, float(Bz), float(Bxy), int(maxj))
    rmax = np.abs(f["rho"]).max()
    noise = 4e-15 * rmax
    assert np.all(np.abs(r - f["rho"]) <= 0.51 * ulp8(f["rho"]) + noise)
    good = np.abs(f["rho"]) > 1e-5 * rmax
    assert good.sum() > 10000
    assert np.all(np.abs(e - f["eng"])[good] <= 0.51 * ulp8(f["eng"])[good] + 2e-9 * 1e3)
    if "esq" in f.files:
        assert np.all(np.abs(q - f["esq"])[good] <= 0.51 * ulp8(f["esq"])[good] + 2e-9 * 1e6)
    same = sum(tg.fmt_e15_8(a) == tg.fmt_e15_8(b) for a, b in zip(r[good][::7], f["rho"][good][::7]))
    assert same >= 0.98 * len(r[good][::7])                 # the remaining lines differ in the last printed digit
