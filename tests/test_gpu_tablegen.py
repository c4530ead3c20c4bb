"""GPU parity tests of the rho-table generators (csrc/pimc_tablegen.cu) through the C ABI.

Checkers: (1) the reference's OWN golden outputs (tests/golden/tablegen/ref_*.npz: nmv_prop/rho.den010_*, nmv_prop/log,
symtop_prop/rho.den0{00,10}_*, the two .rot files of examples/), (2) the quad-precision oracle's outputs for a small
asymmetric top and for Wigner d matrices (oracle_asym_small.npz; the oracle itself is run again where it is fast).

Bars: linear rotor -- BYTE-identical .rot files.  Tops -- the device evaluates the same sums in a factorised order, and
the reference's tables carry rounding noise of ~1e-15 max|rho| from their alternating sums: half a unit of the last
printed digit (E15.8) + that noise against the golden files; 1e-12 max|rho| against the oracle's doubles.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "tablegen")


def ulp8(g):
    g = np.abs(np.asarray(g, dtype=float))
    return 10.0 ** (np.floor(np.log10(np.maximum(g, 1e-300))) + 1 - 8)


@pytest.fixture(scope="module")
def gpu(pkg):
    return pkg.gpu


@pytest.fixture(scope="module")
def tg():
    from oracle import tablegen_py
    return tablegen_py


def test_wigner_d_recurrence_vs_quad_precision_sum(gpu):
    """FP64 upward recurrence in j against Zare's sum evaluated in real*16 (wigd, asymrho.f:1006-1039)"""
    f = np.load(os.path.join(GOLD, "oracle_asym_small.npz"))
    for maxj, n in ((24, 3), (66, 2)):
        for it in range(n):
            th = float(f[f"wigd_j{maxj}_{it}_theta"])
            d = gpu.gen_wigner_d(maxj, th)
            js = f[f"wigd_j{maxj}_{it}_js"]
            ref = f[f"wigd_j{maxj}_{it}"]
            err = np.abs(d[js] - ref).max()
            assert err < 2e-13, (maxj, th, err)
    # closed forms at the poles: d(0) = delta_mk, d(pi) = (-1)^(j-k) delta_{m,-k} in this convention
    maxj = 30
    d0, dpi = gpu.gen_wigner_d(maxj, 0.0), gpu.gen_wigner_d(maxj, np.pi)
    for j in (0, 1, 7, 30):
        w = np.arange(-j, j + 1) + maxj
        assert np.abs(d0[j][np.ix_(w, w)] - np.eye(2 * j + 1)).max() < 1e-13
        anti = np.abs(dpi[j][np.ix_(w, w)])
        assert np.abs(anti - np.eye(2 * j + 1)[::-1]).max() < 1e-10
    # unitarity of every d^j at a generic angle
    d = gpu.gen_wigner_d(maxj, 1.234)
    for j in range(maxj + 1):
        w = np.arange(-j, j + 1) + maxj
        m = d[j][np.ix_(w, w)]
        assert np.abs(m @ m.T - np.eye(2 * j + 1)).max() < 5e-13


def test_wigner_d_large_j_stays_unitary(gpu):
    """the carried exponent keeps the recurrence alive where the start value underflows FP64 (theta near 0 and pi, j up to 200)"""
    maxj = 200
    for th in (0.02, 1.0, np.pi - 0.02):
        d = gpu.gen_wigner_d(maxj, th)
        for j in (90, 150, 200):
            w = np.arange(-j, j + 1) + maxj
            m = d[j][np.ix_(w, w)]
            assert np.abs(m @ m.T - np.eye(2 * j + 1)).max() < 5e-11, (th, j)
            assert abs(np.trace(m) - np.sin((j + 0.5) * th) / np.sin(0.5 * th)) < 1e-9 * (2 * j + 1)      # character of the rotation


def test_wigner_d_against_live_oracle(gpu, tg):
    maxj, th = 12, 2.2
    d = gpu.gen_wigner_d(maxj, th)
    for j in (0, 1, 5, 12):
        for m in range(-j, j + 1):
            for k in range(-j, j + 1):
                assert abs(d[j, m + maxj, k + maxj] - tg.wigd(j, m, k, th)) < 1e-14


def test_asymrho_reference_golden_plane(gpu):
    """asymrho.x 0.37 128 -1 10 10 0.6666525 0.2306476 0.1769383 66 (nmv_prop/a-run) against nmv_prop/rho.den010_*"""
    f = np.load(os.path.join(GOLD, "ref_asymrho_den010.npz"))
    T, ns, io, i0, i1, A, B, C, maxj = f["args"]
    r, e, q, info = gpu.gen_asymrho(float(T), int(ns), int(io), int(i0), int(i1), float(A), float(B), float(C), int(maxj))
    r, e, q = r[0], e[0], q[0]
    rmax = np.abs(f["rho"]).max()
    noise = 1e-13 * rmax                                     # the reference's own rounding noise (values down to -1e-14 in the file)
    assert np.all(np.abs(r - f["rho"]) <= 0.51 * ulp8(f["rho"]) + noise)
    good = np.abs(f["rho"]) > 1e-4 * rmax
    assert good.sum() > 2000
    assert np.all(np.abs(e - f["eng"])[good] <= 0.51 * ulp8(f["eng"])[good] + 1e-8 * np.abs(f["eng"])[good])
    assert np.all(np.abs(q - f["esq"])[good] <= 0.51 * ulp8(f["esq"])[good] + 1e-8 * np.abs(f["esq"])[good])
    txt = [gpu.format_e15_8(v) for v in r[good]]
    ref = [gpu.format_e15_8(v) for v in f["rho"][good]]
    same = sum(a == b for a, b in zip(txt, ref))
    assert same >= 0.97 * len(ref), (same, len(ref))         # identical text; the rest differ by one unit of the last digit
    # log lines "AT BETA" / "AT TAU" (F12.6)
    i = info
    mine = np.array([[i[0], i[1], i[1] / 0.6950356, i[2]], [i[3], i[4], i[4] / 0.6950356, i[5]], [i[6], i[7], i[7] / 0.6950356, i[8]]])
    assert np.all(np.abs(mine - f["at_beta"]) <= 0.5000001e-6)
    mine_tau = np.array([[i[9], i[10], i[10] / 0.6950356], [i[11], i[12], i[12] / 0.6950356], [i[13], i[14], i[14] / 0.6950356]])
    assert np.all(np.abs(mine_tau - f["at_tau"]) <= 0.51e-5 * np.maximum(1.0, np.abs(f["at_tau"])))
    # the plane obeys the symmetries the reference imposes by copying
    assert np.array_equal(r, r.T) and np.array_equal(e, e.T)


@pytest.mark.parametrize("io", [-1, 0, 1])
def test_asymrho_small_top_vs_oracle_planes(gpu, io):
    """whole theta planes (direct region + symmetry fill, every 5th degree) of a water-like top, all three iodevn"""
    f = np.load(os.path.join(GOLD, "oracle_asym_small.npz"))
    T, ns, A, B, C, maxj = f["args"]
    r, e, q, info = gpu.gen_asymrho(float(T), int(ns), io, 0, 180, float(A), float(B), float(C), int(maxj))
    assert np.allclose(info, f[f"io{io}_info"], rtol=1e-12, atol=0)
    for ith in (0, 37, 90, 180):
        ref = f[f"io{io}_th{ith}"]
        rmax = np.abs(ref[0]).max()
        got = np.stack([r[ith], e[ith], q[ith]])[:, ::5, ::5]
        assert np.abs(got[0] - ref[0]).max() <= 1e-12 * rmax, (io, ith)
        good = np.abs(ref[0]) > 1e-6 * rmax
        assert good.sum() > 100
        assert np.all(np.abs(got[1] - ref[1])[good] <= 1e-9 * np.abs(ref[1])[good] + 1e-9)
        assert np.all(np.abs(got[2] - ref[2])[good] <= 1e-9 * np.abs(ref[2])[good] + 1e-7)


def test_asymrho_live_oracle_points_and_sum_rules(gpu, tg):
    T, ns, A, B, C, maxj = 5.0, 4, 9.0, 3.0, 2.0, 14
    r, e, q, info = gpu.gen_asymrho(T, ns, -1, 0, 180, A, B, C, maxj)
    o = tg.AsymRho(T, ns, -1, A, B, C, maxj)
    rmax, emax, qmax = np.abs(r).max(), np.abs(e * r).max(), np.abs(q * r).max()
    rng = np.random.default_rng(7)
    for _ in range(40):
        ith, iphi = int(rng.integers(0, 181)), int(rng.integers(0, 361))
        ichi = int(rng.integers(0, tg.maxchi(iphi) + 1))
        v = o.point(ith, iphi, ichi)
        assert abs(r[ith, iphi, ichi] - v[0]) <= 1e-12 * rmax
        # the estimators are ratios: compare the numerators E*rho, E2*rho, which carry the same absolute rounding noise as rho
        assert abs(e[ith, iphi, ichi] * r[ith, iphi, ichi] - v[1] * v[0]) <= 1e-12 * emax
        assert abs(q[ith, iphi, ichi] * r[ith, iphi, ichi] - v[2] * v[0]) <= 1e-12 * qmax
    # size-independent properties of the full 181 x 361 x 361 table:
    # rho(identity) = Z(tau)/8pi^2 and E(identity) = <E> at tau (sum over m of c_m^2 = 1)
    assert abs(r[0, 0, 0] * 8 * np.pi ** 2 - info[13]) <= 1e-12 * info[13]
    assert abs(e[0, 0, 0] - info[14]) <= 1e-11 * abs(info[14])
    # at theta = 0 the rotation depends on phi + chi only
    for (a, b) in ((10, 20), (100, 3), (200, 100)):
        assert abs(r[0, a, b] - r[0, a + b, 0]) <= 1e-12 * rmax
    # phi <-> chi symmetry and the 360-degree period
    assert np.array_equal(r, np.swapaxes(r, 1, 2))
    assert np.abs(r[:, 0, :] - r[:, 360, :]).max() <= 1e-12 * rmax
    o.close()


def test_asymrho_errors_mirror_the_fortran_stops(gpu):
    with pytest.raises(gpu.PimcGpuError, match="too large contribution from emax"):
        gpu.gen_asymrho(300.0, 1, -1, 0, 0, 0.6666525, 0.2306476, 0.1769383, 10)
    with pytest.raises(gpu.PimcGpuError, match="iodevn can only be"):
        gpu.gen_asymrho(0.37, 128, 2, 0, 0, 0.6666525, 0.2306476, 0.1769383, 66)
    with pytest.raises(gpu.PimcGpuError, match="876"):
        gpu.gen_asymrho(0.37, 128, -1, 0, 0, 0.6666525, 0.2306476, 0.1769383, 900)
    with pytest.raises(gpu.PimcGpuError, match="pmax too large"):
        gpu.gen_symrho(300.0, 1, 1, 0, 0, 0.5, 0.3, 10)


@pytest.mark.parametrize("ith", [0, 10])
def test_symrho_reference_golden_planes(gpu, ith):
    """symrho.x 0.37 128 1 ith ith 0.5 0.3 66 (symtop_prop/a-run) against symtop_prop/rho.den0{00,10}_*"""
    f = np.load(os.path.join(GOLD, f"ref_symrho_den{ith:03d}.npz"))
    T, ns, kmod, _, _, Bz, Bxy, maxj = f["args"]
    r, e, q, info = gpu.gen_symrho(float(T), int(ns), int(kmod), ith, ith, float(Bz), float(Bxy), int(maxj))
    r, e, q = r[0], e[0], q[0]
    rmax = np.abs(f["rho"]).max()
    assert np.all(np.abs(r - f["rho"]) <= 0.51 * ulp8(f["rho"]) + 1e-14 * rmax)
    good = np.abs(f["rho"]) > 1e-5 * rmax
    assert good.sum() > 10000
    assert np.all(np.abs(e - f["eng"])[good] <= 0.51 * ulp8(f["eng"])[good] + 2e-6)
    if "esq" in f.files:
        assert np.all(np.abs(q - f["esq"])[good] <= 0.51 * ulp8(f["esq"])[good] + 2e-3)
    same = sum(gpu.format_e15_8(a) == gpu.format_e15_8(b) for a, b in zip(r[good][::7], f["rho"][good][::7]))
    assert same >= 0.98 * len(r[good][::7])


def test_symrho_vs_live_oracle_kmod(gpu, tg):
    for kmod, ith in ((1, 25), (3, 90), (2, 180)):
        r, e, q, info = gpu.gen_symrho(2.0, 8, kmod, ith, ith, 5.0, 2.5, 30)
        ro, eo, qo, io = tg.symrho_plane(2.0, 8, kmod, ith, 5.0, 2.5, 30)
        rmax = np.abs(ro).max()
        assert np.abs(r[0] - ro).max() <= 1e-13 * rmax
        good = np.abs(ro) > 1e-6 * rmax
        assert np.all(np.abs(e[0] - eo)[good] <= 1e-9 * np.abs(eo)[good] + 1e-9)
        assert np.all(np.abs(q[0] - qo)[good] <= 1e-9 * np.abs(qo)[good] + 1e-7)
        assert np.allclose(info, io, rtol=1e-13)


@pytest.mark.parametrize("name", ["N2O", "CO2"])
def test_linden_byte_identical_to_reference_rot_files(gpu, name, tmp_path):
    f = np.load(os.path.join(GOLD, f"ref_linden_{name}.npz"))
    T, ns, B, npt, io = f["args"]
    out, info = gpu.gen_linden(float(T), int(ns), float(B), int(npt), int(io))
    p, pref = str(tmp_path / "linden.out"), str(tmp_path / "ref.rot")
    gpu.write_rot(p, out)
    gpu.write_rot(pref, f["table"])                              # text -> float64 -> text is the identity (make_fixtures.py)
    assert open(p, "rb").read() == open(pref, "rb").read()
    first = open(p).readline()
    assert first.startswith(" -1.00000000E+00") and len(first) == 65


def test_linden_bit_identical_to_oracle_all_iodevn(gpu, tg):
    for io in (-1, 0, 1):
        out, info = gpu.gen_linden(1.5, 16, 1.92253, 501, io)
        oo, oi = tg.linden(1.5, 16, 1.92253, 501, io)
        assert np.array_equal(out, oo)
        assert np.allclose(info, oi, rtol=1e-14)


def test_table_files_roundtrip_through_the_e15_8_writer(gpu, tmp_path):
    """planes written like rho.denXXX_rho and concatenated like nmv_prop/compile.x; read back as init_rot3D does"""
    r, e, q, _ = gpu.gen_asymrho(10.0, 2, -1, 3, 4, 27.877, 14.512, 9.285, 10)
    p = str(tmp_path / "X_T10t2.rho")
    gpu.write_e15_8(p, r[0])
    gpu.write_e15_8(p, r[1], append=True)
    back = np.loadtxt(p)
    assert back.size == 2 * 361 * 361
    assert np.all(np.abs(back - r.ravel()) <= 0.51 * ulp8(r.ravel()))
    lines = open(p).read().split("\n")
    assert all(len(l) == 15 for l in lines[:-1]) and lines[-1] == ""
    assert lines[0] == gpu.format_e15_8(r[0, 0, 0])


DRV = os.path.join(ROOT, "moribs-pimc_b200", "driver")


def test_pimc_tables_cli_reproduces_reference_files_and_log(pkg, gpu, tmp_path):
    """the command lines of nmv_prop/a-run and linear_prop/README through the pimc_tables binary"""
    import subprocess
    exe = os.path.join(DRV, "pimc_tables")
    if not os.path.exists(exe):
        pytest.skip("pimc_tables not built")
    # linear rotor: the deck's own CO2_T100t4.rot, byte for byte
    out = subprocess.run([exe, "linden", "100", "4", "0.39021", "3000", "-1", "--out", "CO2_T100t4.rot"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    ref = os.path.join(pkg.configs.DECKS, "CO2_100K_4_4", "CO2_T100t4.rot")
    assert open(tmp_path / "CO2_T100t4.rot", "rb").read() == open(ref, "rb").read()
    assert "lmax= 162" in out.stdout
    # asymmetric top: one theta plane with the reference's argument list; files named and formatted as asymrho.x writes them
    out = subprocess.run([exe, "asymrho", "0.37", "128", "-1", "10", "10", "0.6666525", "0.2306476", "0.1769383", "66"], cwd=tmp_path,
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    # nmv_prop/log, verbatim
    assert "EVEN K:    Z=    1.162081 E=    0.136250 CM-1 E=    0.196034 K Cv=    1.842033 Kb" in out.stdout
    assert "ODD  K:    Z=    0.716842 E=    0.488301 CM-1 E=    0.702555 K Cv=    0.719873 Kb" in out.stdout
    assert "CLASSICAL: Z=    1.878923 E=    0.270564 CM-1 E=    0.389280 K Cv=    1.856125 Kb" in out.stdout
    f = np.load(os.path.join(GOLD, "ref_asymrho_den010.npz"))
    for nm in ("rho", "eng", "esq"):
        lines = open(tmp_path / f"rho.den010_{nm}").read().split("\n")
        assert len(lines) == 361 * 361 + 1 and all(len(l) == 15 for l in lines[:-1])
    mine = np.loadtxt(tmp_path / "rho.den010_rho").reshape(361, 361)
    rmax = np.abs(f["rho"]).max()
    assert np.all(np.abs(mine - f["rho"]) <= 1.01 * ulp8(f["rho"]) + 1e-13 * rmax)
    reg = open(tmp_path / "rho.den010").readline()
    assert reg.startswith("   10    0    0  0.62732329E+01  0.26821019E+01 -0.14376686E+04") or reg.startswith("   10    0    0  0.6273232")
    # symmetric top: symtop_prop/a-run, compared with symtop_prop/rho.den010_rho; --table concatenates like compile.x
    out = subprocess.run([exe, "symrho", "0.37", "128", "1", "10", "10", "0.5", "0.3", "66"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ztau=" in out.stdout, out.stdout + out.stderr
    fs = np.load(os.path.join(GOLD, "ref_symrho_den010.npz"))
    mine = np.loadtxt(tmp_path / "rho.den010_rho").reshape(361, 361)
    assert np.all(np.abs(mine - fs["rho"]) <= 1.01 * ulp8(fs["rho"]) + 1e-14 * np.abs(fs["rho"]).max())
    sub = tmp_path / "full"
    sub.mkdir()
    out = subprocess.run([exe, "symrho", "5", "4", "1", "0", "180", "5.0", "2.5", "20", "--table", "X_T5t4"], cwd=sub, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert os.path.getsize(sub / "X_T5t4.rho") == 181 * 361 * 361 * 16 == os.path.getsize(sub / "X_T5t4.esq")
    cat = b"".join(open(sub / f"rho.den{i:03d}_eng", "rb").read() for i in (0, 1, 180))
    full = open(sub / "X_T5t4.eng", "rb").read()
    n1 = 361 * 361 * 16
    assert full[:2 * n1] == cat[:2 * n1] and full[-n1:] == cat[-n1:]
    # usage / error paths
    bad = subprocess.run([exe, "asymrho", "300", "1", "-1", "0", "0", "0.6666525", "0.2306476", "0.1769383", "10"], cwd=tmp_path, capture_output=True, text=True)
    assert bad.returncode == 1 and "too large contribution from emax" in bad.stdout


def test_cxx_driver_generates_missing_rot_table(pkg, tmp_path):
    """pimc_b200 on the reference's CO2 deck WITHOUT its .rot file: the table is generated on the device from the ROTDENSI
    constant (what linden.x would have written), saved under the reference's file name and used by the run"""
    import shutil
    import subprocess
    drv, exe = os.path.join(DRV, "pimc_b200"), os.path.join(DRV, "pimc_tables")
    if not os.path.exists(drv) or not os.path.exists(exe):
        pytest.skip("driver binaries not built")
    d = os.path.join(pkg.configs.DECKS, "CO2_100K_4_4")
    shutil.copy(os.path.join(d, "CO2_fake.pot"), tmp_path)
    deck = open(os.path.join(d, "qmc.input")).read().replace("NUMBEROFBLOCKS     2000  500", "NUMBEROFBLOCKS     6  2")
    deck = deck.replace("OUTPUTDIR        ./g4/1/", "OUTPUTDIR        ./")
    assert "ROTDENSI 0  -1 0.0 0.6666525 0.1769383 0.2306476 1" in deck
    deck = deck.replace("ROTDENSI 0  -1 0.0 0.6666525 0.1769383 0.2306476 1", "ROTDENSI 0  -1 0.0 0.39021 0.39021 0.39021 1")
    open(tmp_path / "qmc.input", "w").write(deck)
    out = subprocess.run([drv, "--chains", "16"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "generating CO2_T100t4.rot on the device" in out.stdout
    sub = tmp_path / "cli"
    sub.mkdir()
    npt = max(1500, int(np.ceil(10.0 / (0.39021 * 1.4387752224 / (100.0 * 4)))) + 1)     # ten grid points per decay length of rho
    assert f"{npt} points" in out.stdout
    o2 = subprocess.run([exe, "linden", "100", "4", "0.39021", str(npt), "-1"], cwd=sub, capture_output=True, text=True, timeout=120)
    assert o2.returncode == 0
    assert open(tmp_path / "CO2_T100t4.rot", "rb").read() == open(sub / "linden.out", "rb").read()
    vals = np.loadtxt(tmp_path / "CO2_monomer.eng", ndmin=2)
    exact = float(o2.stdout.split("Erot at Beta:")[1].split()[0])                         # linden.f:81, 99.8 K
    assert abs(exact - 99.81) < 0.01
    assert vals.shape[0] == 4 and abs(vals[:, 4].mean() - exact) < 3.0                    # free rotor: no Trotter error


def test_free_asymmetric_top_sampled_with_generated_tables_reproduces_exact_energy(pkg, gpu):
    """End to end: tables generated on the device -> rotational moves + GetRotE3D on the device -> <E_rot> of ONE free
    water-like top equals the exact thermal energy sum_n (2J+1) E_n e^{-beta E_n} / Z that asymrho prints as 'AT BETA ...
    CLASSICAL' (asymrho.f:346-368; SURVEY 8c golden vector 3).  The Noya propagator is exact at every tau, so there is no
    Trotter error for a free rotor and the only deviation is statistical.  At 50 K the exact value (70.24 K) lies 6 % under
    the classical 3/2 kT = 75 K, twenty standard errors of this run.  (Step 0.3: with the deck's 0.15 at 20 K / 16 slices the
    ring of orientations changes its SO(3) homotopy class too rarely and the block means drift -- profiles/free_top_check.py.)"""
    T, Q = 50.0, 8
    A, B, C = pkg.configs.ROT_CONSTANTS["H2O"]
    maxj = gpu.asym_auto_maxj(T, Q, A, B, C)
    r, e, q, info = gpu.gen_asymrho(T, Q, -1, 0, 180, A, B, C, maxj)
    exact = info[7] / 0.6950356                                   # K
    assert 69.0 < exact < 71.0
    cfg = pkg.configs.make_config("C4", P=64, Q=Q, big_tables=False, temperature=T)
    cfg.system.types[0].numb = 1
    cfg.system.types[0].rtstep = 0.3
    cfg.coords, cfg.angles = pkg.configs.cluster_config(cfg.system, 3)
    cfg.tables["rot3d"] = (r.reshape(-1), e.reshape(-1), q.reshape(-1))
    G = gpu.PimcGpu(cfg, nchains=148)
    G.seed((4242,) * 6)
    G.steps(400 * cfg.system.P)
    rows = []
    for b in range(40):
        G.accum_reset()
        for k in range(250):
            G.steps(8, sync=False)
            G.measure()
        G.sync()
        s = G.block_scalars()
        rows.append(s.rot / s.count)
    tot, acc = G.counters()
    G.close()
    rows = np.array(rows)
    mean, err = rows.mean(), rows.std(ddof=1) / np.sqrt(len(rows))
    print(f"free top: <E_rot> = {mean:.4f} +- {err:.4f} K, exact {exact:.4f} K, 3/2 kT = {1.5 * T} K, maxj {maxj}, rotational acceptance {acc[0][2] / max(tot[0][2], 1):.3f}")
    assert abs(mean - exact) < 2.5 * err + 2e-3 * exact
    assert err < 0.01 * exact and abs(mean - 1.5 * T) > 8 * err


def test_asymrho_large_basis_sum_rules(gpu):
    """maxj = 150 and 300 (finer imaginary-time steps need larger bases; the reference allows up to 876): rho(identity) = Z(tau)/8 pi^2
    and E(identity) = <E>(tau) to round-off, phi <-> chi symmetry, finite values"""
    for maxj, Q in ((150, 512), (300, 2048)):
        r, e, q, info = gpu.gen_asymrho(0.37, Q, -1, 0, 1, 0.6666525, 0.2306476, 0.1769383, maxj)
        assert abs(r[0, 0, 0] * 8 * np.pi ** 2 - info[13]) <= 1e-12 * info[13]
        assert abs(e[0, 0, 0] - info[14]) <= 1e-11 * abs(info[14])
        assert np.isfinite(r).all() and np.isfinite(e).all() and np.isfinite(q).all()
        assert np.array_equal(r, np.swapaxes(r, 1, 2))
