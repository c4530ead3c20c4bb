"""GPU parity tests for SURVEY section 8 rows a18-a20: area / exchange-length estimators, the symmetry operations of
MCGetAverage (Reflect_MF_*, RotSymConfig) and the rattle-and-shake rotational propagator (RotDenType 1), all through
the C ABI against the CPU oracle (which is pinned bit-exactly on the reference's own objects for the same calls).
"""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SMALL = dict(C5=dict(P=32, Q=8, nsolv=6), C4=dict(P=64, Q=32), C3=dict(P=32, Q=8), C1=dict(P=64, Q=16),
             C2=dict(P=32, Q=8, nsolv=5))


def _oracle():
    from oracle import oracle_py as op
    return op


def _rotor_rows(s):
    rows = np.zeros(s.N * s.P, dtype=bool)
    m = s.types[-1]
    for k in range(m.numb):
        a = s.N - m.numb + k
        rows[a * s.P:a * s.P + s.Q] = True
    return rows


def _perm(nb):
    if nb > 3:
        p = np.arange(nb, dtype=np.int32); p[[0, 1, 2]] = [1, 2, 0]
        return p
    return np.roll(np.arange(nb, dtype=np.int32), 1) if nb > 1 else np.zeros(1, dtype=np.int32)


def _close(a, b, rtol=1e-10):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() <= rtol * scale


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C5"])
def test_area_and_exchange_estimators(pkg, name):
    """a18: GetAreaEstimators (linear dopant), GetAreaEstim3D in both frames, GetExchangeLength, on permuted world lines;
    instantaneous sums to 1e-10 of the largest component, block accumulators of two chains."""
    op = _oracle()
    cfg = copy.copy(pkg.configs.make_config(name, **SMALL[name]))
    s = cfg.system
    nb = s.types[0].numb
    cfg.perm = _perm(nb)
    G = pkg.gpu.PimcGpu(cfg, nchains=2)
    O = op.Oracle(cfg)
    a = G.chain_areas(1)
    oa, oi = O.area_estim3d(0)
    assert _close(a["sff_area"], oa) and _close(a["sff_inert"], oi)
    top = s.types[-1].molecule == 2
    if top:
        ma, mi = O.area_estim3d(1)
        assert _close(a["mff_area"], ma) and _close(a["mff_inert"], mi)
    else:
        ol = O.area_estimators()
        assert _close(a["lin"][:2], ol[:2]) and _close(a["lin"][2:], ol[2:])
    G.accum_reset()
    G.seed((12345,) * 6)
    G.measure()
    acc, lay = G.accum_download()
    A = acc[lay["area"]:lay["area"] + 40]
    tri = lambda v: np.array([v[i] * v[j] for i in range(3) for j in range(i + 1)])
    assert _close(A[6:12], 2 * tri(oa)) and _close(A[12:21], 2 * oi / s.P)
    if top:
        assert _close(A[21:27], 2 * tri(ma)) and _close(A[27:36], 2 * mi / s.P)
        assert not A[0:6].any()
    else:
        assert _close(A[0:2], 2 * ol[:2]) and _close(A[2:4], 2 * ol[:2] ** 2) and _close(A[4:6], 2 * ol[2:] / s.P)
        assert not A[21:36].any()
    pl = acc[lay["ploops"]:lay["ploops"] + nb]
    assert np.array_equal(pl, 2 * O.exchange_length())
    assert pl.sum() == 2 * (nb - 2 if nb > 3 else 1)
    G.close()


@pytest.mark.parametrize("name", ["C1", "C3", "C4", "C5"])
def test_symmetry_operations(pkg, name):
    """a19: explicit Reflect_MF_XZ/YZ/XY and RotSymConfig on the device against the oracle, then the random version
    driven by the chain's miscellaneous MRG32k3a stream (through pimcgpu_measure) against the oracle's replay."""
    op = _oracle()
    cfg = copy.copy(pkg.configs.make_config(name, **SMALL[name]))
    s = cfg.system = copy.copy(cfg.system)
    top = s.types[-1].molecule == 2
    nm = s.types[-1].numb
    s.reflect = (1, 1, 1) if top else (0, 0, 0)
    s.rotsym = 2 if top else 1
    G = pkg.gpu.PimcGpu(cfg, nchains=2)
    O = op.Oracle(cfg)
    rows = _rotor_rows(s)
    seq = [(1, 0, 0, -1), (0, 1, 0, -1), (0, 0, 1, -1), (1, 1, 0, 0), (1, 1, 1, nm - 1)] if top else [(0, 0, 0, 0)]
    for ops in seq:
        G.symmetry_ops([(0, 0, 0, -1), ops])            # chain 0 untouched, chain 1 operated on
        for plane in range(3):
            if ops[plane]:
                O.reflect(plane)
        if ops[3] >= 0:
            O.rotsym((ops[3] + 0.5) / nm, max(1, s.rotsym))
        cg, ag, csg = G.download(1)
        co, ao, cso = O.get_state()
        assert np.abs(cg - co).max() < 1e-12
        assert np.abs(ag[:, rows] - ao[:, rows]).max() < 1e-9, ops
        assert np.abs(csg[:, rows] - cso[:, rows]).max() < 1e-12
    c0, a0, _ = G.download(0)
    assert np.array_equal(c0, cfg.coords) and np.array_equal(a0[:, rows], np.asarray(cfg.angles)[:, rows])
    e, eo = G.chain_energies(1), (O.get_kin(), O.get_pot(0))
    assert abs(e["kin"] - eo[0]) <= 1e-10 * abs(eo[0]) and abs(e["pot"] - eo[1]) <= 1e-9 * abs(eo[1])
    # random application: same stream, same order of draws
    seed = (101, 102, 103, 104, 105, 106)
    G.upload(-1, cfg.coords, cfg.angles, cfg.perm)
    O.set_state(cfg.coords, cfg.angles, cfg.perm)
    G.seed(seed); O.sched_seed(seed, 1)
    nops = 0
    for k in range(12):
        before = O.get_state()[1].copy()
        G.measure()
        O.sched_symmetry(s.reflect[0], s.reflect[1], s.reflect[2], 1 if s.rotsym else 0, max(1, s.rotsym))
        nops += int(not np.array_equal(before, O.get_state()[1]))
    cg, ag, csg = G.download(1)
    co, ao, cso = O.get_state()
    assert nops >= 1
    assert np.abs(cg - co).max() < 1e-12 and np.abs(ag[:, rows] - ao[:, rows]).max() < 1e-8
    # moves after the operations still follow the replay (cached rotor potentials were invalidated)
    G.steps(s.P + 1); O.sched_run(0, s.P + 1)
    cg, ag, _ = G.download(1)
    co, ao, _ = O.get_state()
    assert np.abs(cg - co).max() < 1e-8 and np.abs(ag[:, rows] - ao[:, rows]).max() < 1e-8
    G.close()


@pytest.mark.parametrize("name", ["C4", "C5", "C3"])
def test_rattle_and_shake_propagator(pkg, name):
    """a20: RotDenType 1 (rsrot_ / rsline_ instead of the density tables): rotational estimators and the trajectory of the
    device schedule against the oracle replay with the log-space acceptance of mc_piqmc.cc:903-921,1152-1178."""
    op = _oracle()
    cfg = copy.copy(pkg.configs.make_config(name, big_tables=(name == "C3"), **SMALL[name]))
    s = cfg.system = copy.copy(cfg.system)
    cfg.tables = dict(cfg.tables)
    s.rotden_type, s.rot_odevn, s.rnratio = 1, -1, 1
    top = s.types[-1].molecule == 2
    if top:
        s.x_rot, s.y_rot, s.z_rot = (27.8806, 14.5216, 9.2778) if name == "C4" else (2.02736, 0.34417, 0.29353)
        z = np.zeros(pkg.configs.SIZE_ROTDEN)
        cfg.tables["rot3d"] = (z, z, z)      # GetRotE3D still reads esq from the (unloaded, zero) tables in this mode
    else:
        s.x_rot = 0.419
    G = pkg.gpu.PimcGpu(cfg, nchains=2)
    O = op.Oracle(cfg)
    e = G.chain_energies(1)
    srot, esq, eterm = O.get_rot_energy()
    assert abs(e["rot"] - srot) <= 1e-10 * abs(srot)
    assert abs(e["erotsq"] - esq) <= 1e-10 * max(abs(esq), 1.0) and abs(e["eterm"] - eterm) <= 1e-10 * max(abs(eterm), 1.0)
    seed = (9, 8, 7, 6, 5, 4)
    G.seed(seed); O.sched_seed(seed, 1)
    n = 2 * s.P + 1
    G.steps(n); O.sched_run(0, n)
    cg, ag, _ = G.download(1)
    co, ao, _ = O.get_state()
    rows = _rotor_rows(s)
    assert np.abs(cg - co).max() < 1e-9 and np.abs(ag[:, rows] - ao[:, rows]).max() < 1e-9
    gt, ga = G.counters(); ot, oa = O.counters()
    assert np.array_equal(gt, 2 * ot) and oa[len(s.types) - 1, 2] > 0
    e = G.chain_energies(1)
    srot, esq, eterm = O.get_rot_energy()
    assert abs(e["rot"] - srot) <= 1e-9 * abs(srot)
    G.close()


def test_rcf_legendre_rows_count_time_origins_below_plone(pkg):
    """Rows 1..9 of the block array _rcf: GetRCF (mc_estim.cc:1127-1137) adds pleg = 1 only where n(t0).n(t0+t) < PLONE
    (the Legendre call is commented out, so the `if` governs the accumulation).  The device keeps that count per t in the
    "rcfcnt" region; check it against the downloaded orientations of every chain."""
    cfg = pkg.configs.make_config("C5", P=32, Q=8, nsolv=6)
    s = cfg.system
    G = pkg.gpu.PimcGpu(cfg, nchains=3)
    G.seed((5, 6, 7, 8, 9, 10))
    G.steps(3 * s.P)
    G.accum_reset()
    G.measure()
    G.sync()
    acc, lay = G.accum_download()
    off = G.L.pimcgpu_accum_offset(b"rcfcnt")
    assert off > 0
    Q, P = s.Q, s.P
    rot0 = sum(t.numb for t in s.types if t.molecule == 0) * P          # first rotor's slices
    want = np.zeros(Q)
    want0 = np.zeros(Q)
    for c in range(3):
        _, _, cs = G.download(c)
        n = cs[:, rot0:rot0 + Q]
        for t0 in range(Q):
            for t in range(Q):
                p0 = float(n[:, t0] @ n[:, (t0 + t) % Q])
                want0[t] += p0
                want[t] += 1.0 if p0 < 0.9999999 else 0.0
    assert np.array_equal(acc[off:off + Q], want)
    assert want[0] == 0 and want[1:].sum() > 0                          # n.n = 1 at t = 0 never counts
    assert np.abs(acc[lay["rcf"]:lay["rcf"] + Q] - want0).max() < 1e-10 * Q
    G.close()
