"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on seeded inputs.

Bars (BASELINE.json north_star): table-index selection bit-exact; per-slice potential / kinetic /
rotational energies within 1e-10 relative; trajectories of the device schedule equal to the
oracle's replay of the same schedule (same MRG32k3a streams) to round-off.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-10          # north_star: "match the reference to 1e-10 relative"

SMALL = dict(C5=dict(P=32, Q=8, nsolv=6), C4=dict(P=64, Q=32), C3=dict(P=32, Q=8), C1=dict(P=64, Q=16),
             C2=dict(P=32, Q=8, nsolv=3))


def _oracle():
    from oracle import oracle_py as op
    return op


def relerr(a, b, floor=1e-300):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


@pytest.fixture(scope="module")
def big_tables(pkg):
    """The 181x361x361 synthetic density-matrix tables and a 3-D potential, shared by the top cases."""
    return {}


def make(pkg, name, cache={}, **kw):
    key = (name, tuple(sorted(kw.items())))
    if key not in cache:
        cache[key] = pkg.configs.make_config(name, **(kw or SMALL[name]))
    return cache[key]


# ---------------------------------------------------------------------------------------------------
def test_mrg32k3a_streams_bit_exact(pkg):
    """K13: device integer MRG32k3a == RngStream::U01 (oracle restates rngstream.cc in its double arithmetic)."""
    op = _oracle()
    cfg = make(pkg, "C5")
    G = pkg.gpu.PimcGpu(cfg, nchains=3, chain_offset=5)
    seed = (12345, 12345, 12345, 12345, 12345, 12345)
    G.seed(seed)
    S = cfg.system.P + cfg.system.Q + 8
    for local in (0, 1, S - 1, S, 2 * S + 7):
        stream = 5 * S + local
        d = G.rng_draws(stream, 4000)
        o = op.mrg_draws(seed, stream, 1, 4000)[0]
        assert np.array_equal(d, o), f"stream {stream} differs"
    seed2 = (4294967086, 2, 3, 4294944442, 5, 6)
    G.seed(seed2)
    d = G.rng_draws(5 * S + 3, 1000)
    assert np.array_equal(d, op.mrg_draws(seed2, 5 * S + 3, 1, 1000)[0])
    G.close()


def test_leaf_spline_and_bilinear(pkg):
    """K4: SPot1D / LPot2D / SRotDens* values and table indices."""
    op = _oracle()
    cfg = make(pkg, "C5")
    G = pkg.gpu.PimcGpu(cfg)
    O = op.Oracle(cfg)
    rng = np.random.default_rng(7)
    g = cfg.tables["pot1d"][0]
    r = np.r_[rng.uniform(g[0] * 0.5, g[-1] * 1.2, 20000), g[0], g[-1], g[1], g[-2], g[100], np.nextafter(g[100], 0), np.nextafter(g[100], 99)]
    v, k = G.eval_spot1d(r)
    ov = np.array([O.spot1d(x) for x in r])
    assert np.array_equal(k, ov[:, 1].astype(np.int32)), "spline interval index differs"
    assert relerr(v, ov[:, 0]) < RTOL
    r2 = rng.uniform(1.5, 13.0, 20000); c2 = rng.uniform(-1.05, 1.05, 20000)
    v, ir, ic = G.eval_lpot2d(r2, c2)
    ov = np.array([O.lpot2d(a, b) for a, b in zip(r2, c2)])
    assert np.array_equal(ir, ov[:, 1].astype(np.int32)) and np.array_equal(ic, ov[:, 2].astype(np.int32))
    assert relerr(v, ov[:, 0]) < RTOL
    gm = np.r_[rng.uniform(-1.02, 1.02, 5000), -1.0, 1.0]
    for which in range(3):
        v = G.eval_srotdens(gm, which)
        ov = np.array([O.srotdens(x, which) for x in gm])
        assert np.max(np.abs(v - ov) / np.maximum(np.abs(ov), 1e-12 * np.abs(ov).max())) < RTOL
    G.close()


def test_leaf_helium_nonuniform_grid(pkg):
    """SPot1D on the non-uniform Aziz He-He grid: interval indices must equal the reference's bisection search."""
    op = _oracle()
    cfg = make(pkg, "C2")
    G = pkg.gpu.PimcGpu(cfg)
    O = op.Oracle(cfg)
    g = cfg.tables["pot1d"][0]
    rng = np.random.default_rng(3)
    r = np.r_[rng.uniform(g[0] * 0.8, g[-1] * 1.1, 20000), g]
    v, k = G.eval_spot1d(r)
    ov = np.array([O.spot1d(x) for x in r])
    assert np.array_equal(k, ov[:, 1].astype(np.int32))
    assert relerr(v, ov[:, 0]) < RTOL
    G.close()


def test_leaf_rotden_vcord(pkg):
    """K1/K2: rotden_ (relative Euler angles + rho/E/E^2 gather) and vcord_+vcalc; indices bit-exact."""
    op = _oracle()
    cfg = make(pkg, "C1")
    G = pkg.gpu.PimcGpu(cfg)
    O = op.Oracle(cfg)
    rng = np.random.default_rng(11)
    n = 4000
    e1 = np.c_[rng.uniform(0, 2 * np.pi, n), np.arccos(rng.uniform(-1, 1, n)), rng.uniform(0, 2 * np.pi, n)]
    e2 = e1 + 0.15 * rng.standard_normal((n, 3))          # neighbouring slices: small relative rotation
    e2[:, 1] = np.clip(e2[:, 1], 0, np.pi)
    e2[: n // 4] = np.c_[rng.uniform(0, 2 * np.pi, n // 4), np.arccos(rng.uniform(-1, 1, n // 4)), rng.uniform(0, 2 * np.pi, n // 4)]
    rho, erot, esq, idx = G.eval_rotden(e1, e2)
    o = [O.rotden(a, b) for a, b in zip(e1, e2)]
    oidx = np.array([x[3] for x in o])
    same = idx == oidx
    # an index may differ only where the angle in degrees sits within 1e-9 of an integer (ulp of acos)
    assert same.mean() > 0.999, f"{(~same).sum()} of {n} rho-table indices differ"
    orho, oerot, oesq = (np.array([x[i] for x in o]) for i in range(3))
    assert relerr(rho[same], orho[same], 1e-290) < 1e-9      # interpolation weights carry the acos ulp times 57.3*|slope|
    assert relerr(erot[same], oerot[same]) < 1e-9
    assert relerr(esq[same], oesq[same]) < 1e-9
    # identical orientations (relative theta = 0): the reference's extraction is ill-conditioned there (acos at 1,
    # sign of a ~1e-17 sine picks chi or 2pi-chi), so only physical equivalence is required: rho = rho(0,0,0)
    rho0, _, _, _ = G.eval_rotden(e1[:50], e1[:50])
    assert np.max(np.abs(rho0 / cfg.tables["rot3d"][0][0] - 1.0)) < 1e-4
    eul = np.c_[rng.uniform(0, 2 * np.pi, n), np.arccos(rng.uniform(-1, 1, n)), rng.uniform(0, 2 * np.pi, n)]
    rcom = rng.uniform(-1, 1, (n, 3))
    rpt = rcom + rng.uniform(2.2, 9.0, (n, 1)) * _unit(rng, n)
    v, rtc, vidx = G.eval_vcord(eul, rcom, rpt)
    o = [O.vcord(a, b, c) for a, b, c in zip(eul, rcom, rpt)]
    oidx = np.array([x[2] for x in o])
    same = vidx == oidx
    assert same.mean() > 0.999
    ov = np.array([x[0] for x in o]); ortc = np.array([x[1] for x in o])
    assert relerr(v[same], ov[same], 1e-6) < 1e-9
    assert np.max(np.abs(rtc - ortc)) < 1e-9
    G.close()


def _unit(rng, n):
    u = rng.standard_normal((n, 3))
    return u / np.linalg.norm(u, axis=1)[:, None]


def test_leaf_caleng(pkg):
    """K3: TIP4P pair energy."""
    op = _oracle()
    cfg = make(pkg, "C4")
    G = pkg.gpu.PimcGpu(cfg)
    O = op.Oracle(cfg)
    rng = np.random.default_rng(5)
    n = 5000
    c1 = rng.uniform(-1, 1, (n, 3)); c2 = c1 + rng.uniform(2.4, 8.0, (n, 1)) * _unit(rng, n)
    e1 = np.c_[rng.uniform(0, 2 * np.pi, n), np.arccos(rng.uniform(-1, 1, n)), rng.uniform(0, 2 * np.pi, n)]
    e2 = np.c_[rng.uniform(0, 2 * np.pi, n), np.arccos(rng.uniform(-1, 1, n)), rng.uniform(0, 2 * np.pi, n)]
    e = G.eval_caleng(c1, c2, e1, e2)
    o = np.array([O.caleng(a, b, c, d) for a, b, c, d in zip(c1, c2, e1, e2)])
    # the energy is a sum of ten Coulomb/LJ terms of both signs: bound the error by the largest term scale
    scale = np.maximum(np.abs(o), 1.0)
    assert np.max(np.abs(e - o) / scale) < 1e-9
    G.close()


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4", "C5"])
def test_per_slice_energies(pkg, name):
    """K5/K9: PotEnergy(atom,it) for every bead and the kinetic / potential / rotational estimators, RCF, histograms."""
    op = _oracle()
    cfg = make(pkg, name)
    s = cfg.system
    G = pkg.gpu.PimcGpu(cfg, nchains=2)
    O = op.Oracle(cfg)
    pe = G.pot_energy_slice(1)
    po = np.array([[O.pot_energy_it(a, it) for it in range(s.P)] for a in range(s.N)])
    assert np.max(np.abs(pe - po) / np.maximum(np.abs(po), 1e-3)) < 1e-9, name
    e = G.chain_energies(1)
    assert abs(e["kin"] - O.get_kin()) <= RTOL * abs(O.get_kin())
    assert abs(e["pot"] - O.get_pot(0)) <= 1e-9 * abs(O.get_pot(0))
    if s.Q:
        srot, esq, eterm = O.get_rot_energy()
        assert abs(e["rot"] - srot) <= 1e-9 * abs(srot)
        assert abs(e["erotsq"] - esq) <= 1e-9 * abs(esq)
        assert abs(e["eterm"] - eterm) <= 1e-9 * abs(eterm)
        assert np.max(np.abs(G.chain_rcf(0) - O.get_rcf())) < 1e-10 * s.Q
    # block accumulators: two chains, one measurement each (the decks of C1-C3 enable the symmetry operations, which
    # follow the estimators inside pimcgpu_measure and draw from the seeded miscellaneous stream)
    G.accum_reset()
    G.seed((12345,) * 6)
    G.measure()
    acc, lay = G.accum_download()
    O.reset_hist()
    spot = O.get_pot(1)
    if s.Q:
        O.get_rot_energy()
    h = O.get_hist()
    assert acc[0] == 2.0
    assert abs(acc[2] - 2 * spot) <= 1e-9 * abs(2 * spot)
    assert np.array_equal(acc[lay["gr1d"]:lay["gr1d"] + 300], 2 * h["gr1d"])
    assert np.array_equal(acc[lay["gr2d"]:lay["gr2d"] + 15000], 2 * h["gr2d"])
    if lay["gr3d"] >= 0:
        g3 = acc[lay["gr3d"]:lay["gr3d"] + 1500000]
        assert g3.sum() == 2 * h["gr3d_atoms"].sum()
        assert np.abs(g3 - 2 * h["gr3d_atoms"]).sum() <= 4          # a bead on a bin edge may move by an ulp
    rb = acc[lay["relbins"]:lay["relbins"] + 250]
    if s.Q and s.types[-1].molecule == 2:
        assert np.abs(rb - 2 * np.r_[h["relthe"], h["relphi"], h["relchi"]]).sum() <= 4
    G.close()


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4", "C5"])
def test_schedule_trajectory_matches_oracle_replay(pkg, name):
    """K6/K7/K8: the device move schedule against the oracle's CPU replay with the same MRG32k3a streams."""
    op = _oracle()
    cfg = make(pkg, name)
    s = cfg.system
    G = pkg.gpu.PimcGpu(cfg, nchains=3, chain_offset=2)
    seed = (12345, 23456, 34567, 45678, 56789, 67890)
    G.seed(seed)
    O = op.Oracle(cfg)
    O.sched_seed(seed, 2 + 1)                  # compare local chain 1 == global chain 3
    nsteps = 2 * s.P + 3                       # two full passes: molecular moves at time 0 twice, every bisection offset
    G.steps(nsteps)
    O.sched_run(0, nsteps)
    cg, ag, csg = G.download(1)
    co, ao, cso = O.get_state()
    ot, oa = O.counters()
    # counters of ALL chains: every chain does the same number of attempts
    gt, ga = G.counters()
    assert np.array_equal(gt, 3 * ot)
    assert oa.sum() > 0 and ga.sum() > 0
    dc = np.abs(cg - co).max()
    rows = np.zeros(s.N * s.P, dtype=bool)
    if s.Q:
        m = s.types[-1]
        for k in range(m.numb):
            a = s.N - m.numb + k
            rows[a * s.P:a * s.P + s.Q] = True
    da = np.abs(ag[:, rows] - ao[:, rows]).max() if s.Q else 0.0
    assert dc < 1e-9 and da < 1e-9, f"{name}: trajectory deviates (coords {dc:.2e}, angles {da:.2e})"
    # same trajectory again from the same seed: bit-reproducible on the device
    G.close()
    G3 = pkg.gpu.PimcGpu(cfg, nchains=3, chain_offset=2)
    G3.seed(seed)
    G3.steps(nsteps)
    c2, a2, _ = G3.download(1)
    assert np.array_equal(c2, cg) and np.array_equal(a2, ag)
    G3.close()


def test_geometry_independence(pkg):
    """The trajectory must not depend on cluster size, block size or team width."""
    cfg = make(pkg, "C5")
    seed = (12345,) * 6
    ref = None
    for kw in (dict(), dict(ctas_per_chain=1, threads_per_cta=64, team=4), dict(ctas_per_chain=2, threads_per_cta=128, team=32),
               dict(ctas_per_chain=4, threads_per_cta=32, team=8),
               dict(ctas_per_chain=16, threads_per_cta=128, team=64),          # cooperative grid + software chain barrier, two-warp teams
               dict(ctas_per_chain=2, threads_per_cta=256, team=128)):
        G = pkg.gpu.PimcGpu(cfg, nchains=2, **kw)
        G.seed(seed)
        G.steps(cfg.system.P + 5)
        c, a, _ = G.download(1)
        G.close()
        if ref is None:
            ref = (c, a)
        else:
            assert np.abs(c - ref[0]).max() < 1e-9 and np.abs(a - ref[1]).max() < 1e-9


@pytest.mark.parametrize("name", ["C1", "C2", "C3"])
def test_free_running_top_sweeps_equal_the_staged_ones(pkg, name, monkeypatch):
    """A top whose chain lives in one CTA sweeps its rot slices free-running (rot_run_cta, kernel variant 10: a slice's
    proposal and sums overlap its neighbours' decisions) instead of in three CTA-wide stages (rot_sweep_pipe).  Same draws,
    same sums, same order of decisions seen by every slice: the trajectories are identical bit for bit, so every statistical
    and oracle comparison made with either path holds for the other."""
    cfg = make(pkg, name)
    s = cfg.system
    out = []
    for staged in (False, True):
        if staged:
            monkeypatch.setenv("PIMC_NO_ROT_RUN", "1")
        else:
            monkeypatch.delenv("PIMC_NO_ROT_RUN", raising=False)
        G = pkg.gpu.PimcGpu(cfg, nchains=3, ctas_per_chain=1)
        G.seed((11, 12, 13, 14, 15, 16))
        G.steps(2 * s.P + 3)                       # translational sweeps in between: the stretches restart
        G.steps(5)                                 # a second launch picks the counters up again
        st = [G.download(c)[:2] for c in range(3)]
        out.append((st, G.counters(), G.geometry()))
        G.close()
    assert out[0][2]["kind"] == 10 and out[1][2]["kind"] == 2 and out[0][2]["ctas_per_chain"] == 1      # variant 10 = top, free-running
    for (c0, a0), (c1, a1) in zip(out[0][0], out[1][0]):
        assert np.array_equal(c0, c1) and np.array_equal(a0, a1)
    assert np.array_equal(out[0][1][0], out[1][1][0]) and np.array_equal(out[0][1][1], out[1][1][1])
    assert out[0][1][1][-1][2] > 0                  # rotations were accepted


def test_permuted_world_lines_exchange_paths(pkg):
    """BOSE species with a non-identity permutation: bisection segments that cross beta continue on PIndex[atom],
    whole-path moves shift complete cycles, the kinetic estimator closes each path on its successor (a3, a4, a14)."""
    import copy
    op = _oracle()
    cfg = copy.copy(make(pkg, "C2"))
    cfg.perm = np.array([1, 2, 0], dtype=np.int32)            # one 3-cycle
    s = cfg.system
    G = pkg.gpu.PimcGpu(cfg, nchains=2)
    O = op.Oracle(cfg)
    e = G.chain_energies(0)
    assert abs(e["kin"] - O.get_kin()) <= RTOL * abs(O.get_kin())
    seed = (777, 778, 779, 780, 781, 782)
    G.seed(seed); O.sched_seed(seed, 1)
    n = 2 * s.P + 1
    G.steps(n); O.sched_run(0, n)
    cg, ag, _ = G.download(1)
    co, ao, _ = O.get_state()
    assert np.abs(cg - co).max() < 1e-9
    ot, oa = O.counters()
    gt, ga = G.counters()
    assert np.array_equal(gt, 2 * ot) and oa[0, 1] > 0 and oa[0, 0] >= 0
    assert ot[0, 0] == 3                                      # three passes started (t = 0, P, 2P); one 3-cycle -> ONE whole-path move each
    e = G.chain_energies(1)
    assert abs(e["kin"] - O.get_kin()) <= 1e-9 * abs(O.get_kin())
    G.close()


def test_odd_rot_slices_and_ragged_segments(pkg):
    """Q odd (third rotational phase for the last slice) and P not a multiple of the segment length."""
    op = _oracle()
    cfg = pkg.configs.make_config("C5", P=36, Q=9, nsolv=5)
    s = cfg.system
    G = pkg.gpu.PimcGpu(cfg, nchains=1)
    O = op.Oracle(cfg)
    seed = (31, 32, 33, 34, 35, 36)
    G.seed(seed); O.sched_seed(seed, 0)
    n = 3 * s.P
    G.steps(n); O.sched_run(0, n)
    cg, ag, _ = G.download(0)
    co, ao, _ = O.get_state()
    rows = np.zeros(s.N * s.P, dtype=bool); rows[(s.N - 1) * s.P:(s.N - 1) * s.P + s.Q] = True
    assert np.abs(cg - co).max() < 1e-9 and np.abs(ag[:, rows] - ao[:, rows]).max() < 1e-9
    gt, _ = G.counters(); ot, _ = O.counters()
    assert np.array_equal(gt, ot)
    G.close()


def test_minimum_image_and_error_paths(pkg):
    """MINIMAGE pair distances; init-time errors are reported, not fatal."""
    op = _oracle()
    cfg = pkg.configs.make_config("C2", P=32, Q=8, nsolv=4, big_tables=True)
    cfg.system.minimage = 1
    cfg.system.density = 0.004
    G = pkg.gpu.PimcGpu(cfg, nchains=1)
    O = op.Oracle(cfg)
    pe = G.pot_energy_slice(0)
    po = np.array([[O.pot_energy_it(a, it) for it in range(cfg.system.P)] for a in range(cfg.system.N)])
    assert np.max(np.abs(pe - po) / np.maximum(np.abs(po), 1e-3)) < 1e-9
    G.close()
    bad = pkg.configs.make_config("C5", P=8, Q=4, nsolv=2)          # segment 2^3 is not smaller than P
    with pytest.raises(pkg.gpu.PimcGpuError, match="segment size"):
        pkg.gpu.PimcGpu(bad)
    rs = pkg.configs.make_config("C4", P=64, Q=32)
    rs.system.rotden_type = 2
    with pytest.raises(pkg.gpu.PimcGpuError, match="RotDenType"):
        pkg.gpu.PimcGpu(rs)
    G = pkg.gpu.PimcGpu(make(pkg, "C5"))
    with pytest.raises(pkg.gpu.PimcGpuError, match="seed"):
        G.steps(1)                                                   # stepping before pimcgpu_seed
    with pytest.raises(pkg.gpu.PimcGpuError):
        G.seed((0, 0, 0, 1, 2, 3))                                   # CheckSeed: first triple all zero
    G.close()


def test_batched_state_transfer_equals_per_chain_calls(pkg):
    """pimcgpu_upload_states / pimcgpu_download_states move every chain in one call (one copy + one transposing kernel);
    the device state and the trajectory that follows must be those of the per-chain entry points."""
    cfg = make(pkg, "C2")
    s = cfg.system
    n, C = s.N * s.P, 5
    rng = np.random.default_rng(3)
    nb = s.types[0].numb
    coords = np.stack([cfg.coords + 0.01 * rng.standard_normal(cfg.coords.shape) for _ in range(C)])
    angles = np.stack([cfg.angles.copy() for _ in range(C)])
    rot0 = nb * s.P
    for c in range(C):
        angles[c, 0, rot0:rot0 + s.Q] = rng.uniform(0, 2 * np.pi, s.Q)
        angles[c, 1, rot0:rot0 + s.Q] = rng.uniform(-0.9, 0.9, s.Q)
        angles[c, 2, rot0:rot0 + s.Q] = rng.uniform(0, 2 * np.pi, s.Q)
    perms = np.stack([rng.permutation(nb) for _ in range(C)]).astype(np.int32)
    results = []
    for batched in (True, False):
        G = pkg.gpu.PimcGpu(cfg, nchains=C)
        if batched:
            G.upload_all(coords, angles, perms)
        else:
            for c in range(C):
                G.upload(c, coords[c], angles[c], perms[c])
        # what was uploaded comes back, through either download path
        ca, aa, sa = (np.zeros((C, 3, n)) for _ in range(3))
        G.download_all_into(ca, aa, sa)
        for c in range(C):
            c1, a1, s1 = G.download(c)
            assert np.array_equal(c1, ca[c]) and np.array_equal(a1, aa[c]) and np.array_equal(s1, sa[c])
            assert np.array_equal(c1, coords[c])
            assert np.array_equal(a1[:, rot0:rot0 + s.Q], angles[c][:, rot0:rot0 + s.Q])
            assert np.array_equal(G.download_perm(c), perms[c])
        G.seed((21, 22, 23, 24, 25, 26))
        G.steps(s.P + 7)
        G.download_all_into(ca, aa, sa)
        results.append((ca.copy(), aa.copy(), G.counters()))
        G.close()
    assert np.array_equal(results[0][0], results[1][0]) and np.array_equal(results[0][1], results[1][1])
    assert np.array_equal(results[0][2][0], results[1][2][0]) and np.array_equal(results[0][2][1], results[1][2][1])
    assert results[0][2][1].sum() > 0


def test_full_size_c5_energies_and_checksum_properties(pkg):
    """BASELINE configs[4] at its full size (N = 101, P = 1024, Q = 128) after a stretch of sampling: the estimators of the
    device against the oracle on the downloaded configuration, and the size-independent properties
      sum over all beads of PotEnergy(atom, it) = 2 P <V>          (every pair term appears once per partner), 
      histogram counts = number of binned pair terms,
      the state that went through the batched export / import round trip is unchanged."""
    op = _oracle()
    cfg = pkg.configs.make_config("C5")
    s = cfg.system
    assert (s.N, s.P, s.Q) == (101, 1024, 128)
    G = pkg.gpu.PimcGpu(cfg, nchains=2)
    G.seed((9, 8, 7, 6, 5, 4))
    G.steps(2 * 128 + 3)                               # whole-path sweep, three bisection sweeps, 259 rotational sweeps
    c, a, _ = G.download(1)
    O = op.Oracle(cfg)
    O.set_state(c, a)
    e = G.chain_energies(1)
    ko, po = O.get_kin(), O.get_pot(0)
    assert abs(e["kin"] - ko) <= RTOL * abs(ko)
    assert abs(e["pot"] - po) <= 1e-9 * abs(po)
    srot, esq, eterm = O.get_rot_energy()
    assert abs(e["rot"] - srot) <= 1e-9 * abs(srot) and abs(e["erotsq"] - esq) <= 1e-9 * abs(esq)
    pe = G.pot_energy_slice(1)                          # [N][P]
    assert abs(pe.sum() / (2.0 * s.P) - e["pot"]) <= 1e-10 * abs(e["pot"])
    for (atom, it) in ((0, 0), (57, 511), (100, 1023), (100, 64)):
        assert abs(pe[atom, it] - O.pot_energy_it(atom, it)) <= 1e-9 * max(1e-3, abs(pe[atom, it]))
    G.accum_reset()
    G.measure()
    acc, lay = G.accum_download()
    assert acc[0] == 2.0
    npair_aa = 100 * 99 // 2
    g1 = acc[lay["gr1d"]:lay["gr1d"] + 300]
    g2 = acc[lay["gr2d"]:lay["gr2d"] + 15000]
    assert g1.sum() <= 2 * npair_aa * s.P and g1.sum() > 0.5 * 2 * npair_aa * s.P       # pairs inside 15 Angstrom
    assert g2.sum() <= 2 * 100 * s.P and g2.sum() > 0.9 * 2 * 100 * s.P
    ca, aa = np.zeros((2, 3, s.N * s.P)), np.zeros((2, 3, s.N * s.P))
    G.download_all_into(ca, aa)
    G.upload_all(ca, aa)
    c2, a2, _ = G.download(1)
    assert np.array_equal(c2, c) and np.array_equal(a2, a)
    G.close()


def test_split_phase_transfers_equal_the_synchronous_calls(pkg):
    """pimcgpu_upload_states_begin/_commit and pimcgpu_download_states_begin/_end (copies on a second stream, overlapping the move
    kernel of the neighbouring steps): two sets of chains alternate on the device exactly as if each had been moved with the
    synchronous calls -- same states back, same trajectories."""
    cfg = make(pkg, "C5")
    s = cfg.system
    n, C = s.N * s.P, 3
    rng = np.random.default_rng(17)
    sets = [np.ascontiguousarray(np.stack([cfg.coords + 0.01 * rng.standard_normal(cfg.coords.shape) for _ in range(C)])) for _ in range(2)]
    angs = [np.ascontiguousarray(np.stack([cfg.angles.copy() for _ in range(C)])) for _ in range(2)]
    results = []
    for split in (False, True):
        G = pkg.gpu.PimcGpu(cfg, nchains=C)
        G.seed((41, 42, 43, 44, 45, 46))
        hc = [x.copy() for x in sets]; ha = [x.copy() for x in angs]
        accs = []
        if split:
            acc = np.zeros(G.accum_layout()["n_total"])
            G.upload_begin(hc[0], ha[0])
            for k in range(6):
                cur, oth = k % 2, 1 - k % 2
                G.upload_commit()
                G.accum_reset()
                G.steps(7, sync=False)
                G.measure()
                G.L.pimcgpu_accum_device_ptr()
                G.accum_download_begin(acc)              # the sums first, the configuration behind them
                if k > 0:
                    G.download_end()
                G.upload_begin(hc[oth], ha[oth])
                G.download_begin(hc[cur], ha[cur])
                G.accum_download_end()
                accs.append(acc.copy())
            G.download_end()
            G.upload_commit()
        else:
            for k in range(6):
                cur = k % 2
                G.upload_all(hc[cur], ha[cur])
                G.accum_reset()
                G.steps(7)
                G.measure()
                accs.append(G.accum_download()[0])
                G.download_rows_into(hc[cur], ha[cur])
        results.append((hc, ha, G.counters(), accs))
        G.close()
    for a, b in zip(results[0][3], results[1][3]):
        assert np.array_equal(a, b) and np.abs(a).sum() > 0
    for a, b in zip(results[0][0], results[1][0]):
        assert np.array_equal(a, b)
    for a, b in zip(results[0][1], results[1][1]):
        assert np.array_equal(a, b)
    assert np.array_equal(results[0][2][0], results[1][2][0]) and np.array_equal(results[0][2][1], results[1][2][1])
    assert not np.array_equal(results[0][0][0], sets[0])
