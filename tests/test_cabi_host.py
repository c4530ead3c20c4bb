"""CPU-only tests: the C-ABI library loads and exports every symbol include/pimcgpu.h declares, fails loudly
without a GPU, and its host-side table preparation agrees with the oracle."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "pimcgpu.h")).read()
    declared = set(re.findall(r"\b(pimcgpu_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = C.CDLL(pkg.gpu.build())
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert declared == set(pkg.gpu.EXPORTS), declared ^ set(pkg.gpu.EXPORTS)


def test_no_cpu_fallback(pkg):
    """Without a device the product path must fail loudly (nonzero status + message), never compute on the CPU."""
    if _has_gpu():
        pytest.skip("a GPU is present")
    cfg = pkg.configs.make_config("C5", P=32, Q=8, nsolv=4)
    with pytest.raises(pkg.gpu.PimcGpuError) as e:
        pkg.gpu.PimcGpu(cfg)
    assert "CUDA" in str(e.value)
    L = pkg.gpu.lib()
    assert L.pimcgpu_steps(C.c_long(1)) != 0 and b"not initialised" in L.pimcgpu_last_error()
    assert L.pimcgpu_measure() != 0
    out = np.zeros(4)
    assert L.pimcgpu_eval_spot1d(4, out.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)), None) != 0


def test_table_generators_fail_loudly_without_gpu_and_formatters_work_on_host(pkg, tmp_path):
    """The generators are device code with no CPU path; the Fortran edit descriptors and file writers are host code and must
    agree with the oracle's formatter (which is pinned on the reference's files)."""
    import subprocess
    from oracle import tablegen_py as tg
    g = pkg.gpu
    for v in (6.2732329, -0.50630561, 0.0, -1437.6686, 9.99999999e-5, 1.5e-120, -3.25e+105, 1.0, 0.099999999999):
        assert g.format_e15_8(v) == tg.fmt_e15_8(v)
        assert g.format_e15_8(v, True) == tg.fmt_e15_8(v, True)[:15]
    vals = np.array([1.0, -2.5e-7, 3.0e10, 0.0])
    p = str(tmp_path / "t.rho")
    g.write_e15_8(p, vals)
    g.write_e15_8(p, vals[:2], append=True)
    assert open(p).read() == "".join(tg.fmt_e15_8(v) + "\n" for v in list(vals) + list(vals[:2]))
    rot = np.arange(8.0).reshape(2, 4) - 3.5
    g.write_rot(str(tmp_path / "t.rot"), rot)
    assert open(tmp_path / "t.rot").readlines() == tg.rot_lines(rot)
    if _has_gpu():
        return
    for call in (lambda: g.gen_linden(0.5, 128, 0.419, 100, -1), lambda: g.gen_asymrho(10.0, 2, -1, 0, 0, 27.9, 14.5, 9.3, 10),
                 lambda: g.gen_symrho(5.0, 4, 1, 0, 0, 5.0, 2.5, 20), lambda: g.gen_wigner_d(4, 0.3)):
        with pytest.raises(g.PimcGpuError, match="no CUDA device"):
            call()
    exe = os.path.join(ROOT, "moribs-pimc_b200", "driver", "pimc_tables")
    if os.path.exists(exe):
        out = subprocess.run([exe, "linden", "0.5", "128", "0.419", "1400", "-1"], cwd=tmp_path, capture_output=True, text=True)
        assert out.returncode == 1 and "no CUDA device" in out.stdout and not os.path.exists(tmp_path / "linden.out")
        assert subprocess.run([exe], capture_output=True, text=True).stdout.startswith("usage: pimc_tables asymrho")


def test_host_spline_setup_matches_oracle(pkg):
    from oracle import oracle_py as op
    L = pkg.gpu.lib()
    for name in ("C2", "C3", "C5"):
        cfg = pkg.configs.make_config(name, P=32, Q=8, nsolv=2, big_tables=False)
        g, v = (np.ascontiguousarray(x) for x in cfg.tables["pot1d"])
        n = len(g)
        y2, auc = np.zeros(n), np.zeros(3)
        assert L.pimcgpu_host_spline(n, op._dp(g), op._dp(v), op._dp(y2), op._dp(auc)) == 0
        O = op.Oracle(pkg.configs.make_config(name, P=32, Q=8, nsolv=2, big_tables=False)) if name == "C5" else None
        if O is None:
            import ctypes
            lib = ctypes.CDLL(op.build_port())
            lib.orc_create.restype = ctypes.c_void_p
            sy = op.system_struct(cfg.system)
            h = ctypes.c_void_p(lib.orc_create(ctypes.byref(sy)))
            lib.orc_set_pot1d(h, n, op._dp(g), op._dp(v))
        else:
            lib, h = O.lib, O.h
        oy2, oauc = np.zeros(n), np.zeros(3)
        lib.orc_get_pot1d_setup(h, op._dp(oy2), op._dp(oauc))
        assert np.allclose(y2, oy2, rtol=1e-12, atol=1e-300)
        assert np.allclose(auc, oauc, rtol=1e-12)
        # bucket table: lut[b] must be a lower bound of the interval index of every x in the bucket
        lut = np.zeros(4 * n, dtype=np.int32); sc = C.c_double()
        nl = L.pimcgpu_host_lut(n, op._dp(g), lut.ctypes.data_as(C.POINTER(C.c_int)), C.byref(sc))
        assert nl == 4 * n
        xs = np.random.default_rng(0).uniform(g[0], g[-1], 20000)
        b = np.clip(((xs - g[0]) * sc.value).astype(int), 0, nl - 1)
        klo = np.clip(np.searchsorted(g, xs, side="right") - 1, 0, n - 2)
        assert np.all(lut[b] <= klo) and np.all(klo - lut[b] <= 2 + (np.diff(g).max() / np.diff(g).min() > 1.5) * n)


def test_host_stream_jump_matches_rngstream(pkg):
    """(A^(2^127))^s seed in exact integer arithmetic == the RngStream constructor chain restated by the oracle."""
    from oracle import oracle_py as op
    L = pkg.gpu.lib()
    lo = C.CDLL(op.build_port())
    for seed in ((12345,) * 6, (1, 2, 3, 4, 5, 6), (4294967086, 7, 8, 4294944442, 9, 10)):
        for s in (0, 1, 2, 1159, 1160, 10 ** 6 + 3, 2 ** 40 + 17):
            st = (C.c_ulong * 6)()
            assert L.pimcgpu_host_stream_state((C.c_ulong * 6)(*seed), C.c_long(s), st) == 0
            o = np.zeros(6)
            lo.orc_mrg_stream_state((C.c_ulong * 6)(*seed), C.c_long(s), op._dp(o))
            assert list(st) == list(o.astype(np.uint64)), (seed, s)


def test_decks_and_bead_update_counts(pkg):
    """The five BASELINE configurations and the bead-update table of SURVEY.md 8(d) / BASELINE.md section 3."""
    want = {"C1": (23552, 1024, 65536), "C2": (69120, 4608, 65536), "C3": (76800, 5120, 262144),
            "C4": (253952, 8192, 16777216), "C5": (723968, 103424, 131072)}
    for name, (bis, mol, rot) in want.items():
        s = pkg.configs.make_config(name, big_tables=False).system
        bu = s.bead_updates_per_pass()
        assert (bu["bisection"], bu["molecular"], bu["rotation"]) == (bis, mol, rot), name
        assert bu["total"] == bis + mol + rot
    c2 = pkg.configs.make_config("C2", big_tables=False)
    assert list(c2.perm) == [2, 1, 3, 7, 6, 0, 5, 4] and c2.coords.shape == (3, 9 * 512)
    c3 = pkg.configs.make_config("C3", big_tables=False)
    assert list(c3.perm) == [0, 2, 3, 1] and c3.system.reflect == (1, 1, 0) and c3.system.rotsym == 2
    # deck writer -> parser round trip (the writer feeds the reference's own parser in oracle/_ref)
    import tempfile
    s = c3.system
    with tempfile.TemporaryDirectory() as d:
        pkg.configs.write_qmc_input(s, os.path.join(d, "qmc.input"))
        t = pkg.configs.parse_qmc_input(os.path.join(d, "qmc.input"))
    assert (t.P, t.Q, t.temperature, [x.numb for x in t.types], [x.levels for x in t.types]) == \
           (s.P, s.Q, s.temperature, [x.numb for x in s.types], [x.levels for x in s.types])
