"""GPU parity tests for the worm algorithm on the device: the world-line masks of
the potential sums, single MCWormMove calls and the full device schedule with exchange sampling, against the CPU
oracle (pinned bit-exactly on the reference's mc_qworm.cc objects by tests/test_oracle.py)."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SMALL = dict(C2=dict(P=32, Q=8, nsolv=5), C3=dict(P=32, Q=8), C5=dict(P=32, Q=8, nsolv=6))
WORM = dict(C2=("He4", 0.13, 8), C3=("H2", 0.35, 8), C5=("H2", 0.35, 6))


def _oracle():
    from oracle import oracle_py as op
    return op


def make(pkg, name):
    cfg = copy.copy(pkg.configs.make_config(name, **SMALL[name]))
    s = cfg.system = copy.copy(cfg.system)
    s.worm = WORM[name]
    s.reflect, s.rotsym = (0, 0, 0), 0
    nb = s.types[0].numb
    c = cfg.coords.reshape(3, s.N, s.P).copy()
    com = c[:, :nb, :].mean(axis=(1, 2), keepdims=True)
    c[:, :nb, :] = com + 0.8 * (c[:, :nb, :] - com)          # a cluster compact enough for exchange
    cfg.coords = np.ascontiguousarray(c.reshape(3, -1))
    return cfg


def rotor_rows(s):
    rows = np.zeros(s.N * s.P, dtype=bool)
    rows[(s.N - 1) * s.P:(s.N - 1) * s.P + s.Q] = True
    return rows


@pytest.mark.parametrize("name", ["C2", "C3", "C5"])
def test_world_line_masks(pkg, name):
    """a22: PotEnergy(atom, it) for every bead with an open worm (gap inside one world line, and across beta on two)."""
    op = _oracle()
    cfg = make(pkg, name)
    s = cfg.system
    nb = s.types[0].numb
    cfg.perm = np.roll(np.arange(nb, dtype=np.int32), 1)
    G = pkg.gpu.PimcGpu(cfg, nchains=2)
    O = op.Oracle(cfg)
    for st in ((1, 5, 11, 1, 1), (1, s.P - 3, 4, 2, int(cfg.perm[2])), (0, 3, 9, 0, 0)):
        G.worm_set(1, st); O.worm_set(st)
        pe = G.pot_energy_slice(1)
        po = np.array([[O.pot_energy_it(a, it) for it in range(s.P)] for a in range(s.N)])
        assert np.max(np.abs(pe - po) / np.maximum(np.abs(po), 1e-3)) < 1e-9, st
        p0 = G.pot_energy_slice(0)                       # chain 0 stays closed
        G.worm_set(1, (0, 0, 0, 0, 0)); O.worm_set((0, 0, 0, 0, 0))
        assert np.array_equal(p0, G.pot_energy_slice(1))
    G.close()


@pytest.mark.parametrize("name", ["C2", "C3", "C5"])
def test_worm_moves_match_oracle(pkg, name):
    """N1: sequences of MCWormMove on the device against the oracle replay drawing from the same MRG32k3a stream."""
    op = _oracle()
    cfg = make(pkg, name)
    s = cfg.system
    nb = s.types[0].numb
    G = pkg.gpu.PimcGpu(cfg, nchains=3, chain_offset=1)
    O = op.Oracle(cfg)
    seed = (4242, 4243, 4244, 4245, 4246, 4247)
    G.seed(seed); O.sched_seed(seed, 1 + 2)          # local chain 2 == global chain 3
    opened = 0
    for k in range(60):
        for _ in range(5):
            G.worm_moves(sync=False); O.worm_op(7, sched_stream=True)
        G.sync()
        st = G.worm_state(2)
        assert st == O.worm_get(), (k, st, O.worm_get())
        opened += st[0]
        cg, _, _ = G.download(2)
        co, _, _ = O.get_state()
        assert np.abs(cg - co).max() < 1e-9, k
        assert np.array_equal(G.download_perm(2), O.get_perm(nb)[0]), k
    t, a, cq = O.worm_counters()
    gt, ga, gq = G.worm_counters()
    assert gt[0] + gt[1] == 3 * (t[0] + t[1])        # every call tries one open-or-close per atom in every chain; the rest depends on the chain's history
    assert a[0] > 0 and a[1] > 0 and a[4] > 0 and a[5] > 0 and opened > 0
    if name == "C2":
        assert a[6] > 0, "no accepted swap exercised"
    G.close()


@pytest.mark.parametrize("name", ["C2", "C5"])
def test_schedule_with_worm(pkg, name):
    """The device schedule with WORM: worm moves every step, path moves of the worm's species only in the Z sector,
    rotor moves with the world-line masks while the worm is open; estimators skip chains in the G sector."""
    op = _oracle()
    cfg = make(pkg, name)
    s = cfg.system
    s.worm = (s.worm[0], 0.003, s.worm[2])       # a small C: both sectors are visited within a few passes
    nb = s.types[0].numb
    G = pkg.gpu.PimcGpu(cfg, nchains=2)
    O = op.Oracle(cfg)
    seed = (77, 78, 79, 80, 81, 82)
    G.seed(seed); O.sched_seed(seed, 1)
    n = 3 * s.P + 2
    G.steps(n); O.sched_run(0, n)
    assert G.worm_state(1) == O.worm_get()
    cg, ag, _ = G.download(1)
    co, ao, _ = O.get_state()
    rows = rotor_rows(s)
    assert np.abs(cg - co).max() < 1e-8 and np.abs(ag[:, rows] - ao[:, rows]).max() < 1e-8
    assert np.array_equal(G.download_perm(1), O.get_perm(nb)[0])
    ot, oa = O.counters(); gt, ga = G.counters()
    wt, wa, _ = O.worm_counters()
    assert wa[0] > 0 and wa[1] > 0 and oa[0, 1] > 0 and oa[0, 0] > 0 and oa[1, 2] > 0
    # estimators: only closed chains count
    G.accum_reset()
    G.measure()
    acc, lay = G.accum_download()
    nclosed = sum(1 - G.worm_state(c)[0] for c in range(2))
    assert acc[0] == nclosed
    G.close()


@pytest.mark.parametrize("name", ["C2", "C5"])
def test_checkpoint_restart_is_exact(pkg, name):
    """N4: a context that loads pimcgpu_checkpoint_save's blob continues bit-identically (beads, angles, permutation
    tables, worm, MRG32k3a streams, rotor-potential cache, step counter), here with exchange sampling switched on."""
    cfg = make(pkg, name)
    s = cfg.system
    s.worm = (s.worm[0], 0.003, s.worm[2])
    nb = s.types[0].numb
    G = pkg.gpu.PimcGpu(cfg, nchains=2, chain_offset=4)
    G.seed((5, 6, 7, 8, 9, 10))
    n1, n2 = s.P + 7, 2 * s.P + 3
    G.steps(n1)
    blob = G.checkpoint_save()
    G.steps(n2)
    ref = [G.download(c) for c in range(2)]
    refw = [G.worm_state(c) for c in range(2)]
    refp = [G.download_perm(c) for c in range(2)]
    G.close()
    G2 = pkg.gpu.PimcGpu(cfg, nchains=2, chain_offset=4)
    G2.checkpoint_load(blob)
    assert G2.L.pimcgpu_step_counter() == n1
    G2.steps(n2)
    for c in range(2):
        got = G2.download(c)
        assert np.array_equal(got[0], ref[c][0]) and np.array_equal(got[1], ref[c][1])
        assert G2.worm_state(c) == refw[c] and np.array_equal(G2.download_perm(c), refp[c])
    # a blob of another system is refused
    G2.close()
    other = pkg.gpu.PimcGpu(cfg, nchains=3, chain_offset=4)
    with pytest.raises(pkg.gpu.PimcGpuError, match="different system"):
        other.checkpoint_load(blob)
    other.close()
