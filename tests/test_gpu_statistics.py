"""Converged-estimator parity (north_star, third check): <K>, <V>, <E_rot> sampled by the CUDA path agree with the
reference's own CPU sampling (oracle/_ref: its unmodified move and estimator code, SPRNG + MRG32k3a streams) within 2 sigma
of the combined statistical error; and the C++ driver writes the reference's .eng format."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REF_SCRIPT = r'''
import sys, json, numpy as np
sys.path.insert(0, %(root)r)
from oracle import oracle_py as op
import ctypes as C
cfgs = op._configs()
cfg = cfgs.make_config(%(name)r, **%(kw)r)
R = op.Ref(cfg)
s = cfg.system
nblocks, per_block, skip = %(nblocks)d, %(per_block)d, %(skip)d
R.lib.ref_run_steps(0, 100 * s.P)                     # equilibrate
out7 = np.zeros(7)
rows = []
t = 100 * s.P
for b in range(nblocks):
    R.lib.ref_reset_block()
    n = 0
    for k in range(per_block // skip):
        R.lib.ref_run_steps(t, skip); t += skip
        R.lib.ref_MCGetAverage(op._dp(out7)); n += 1
    rows.append((out7[:3] / n).tolist())
print("ROWS " + json.dumps(rows))
'''


def reference_blocks(name, kw, nblocks, per_block, skip):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpimcref.so")):
        pytest.skip("oracle/_ref not available")
    code = REF_SCRIPT % dict(root=ROOT, name=name, kw=kw, nblocks=nblocks, per_block=per_block, skip=skip)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    line = [l for l in out.stdout.splitlines() if l.startswith("ROWS ")]
    assert line, out.stdout[-1500:] + out.stderr[-1500:]
    import json
    return np.array(json.loads(line[-1][5:]))


def gpu_blocks(pkg, name, kw, nchains, nblocks, per_block, skip):
    cfg = pkg.configs.make_config(name, **kw)
    G = pkg.gpu.PimcGpu(cfg, nchains=nchains)
    G.seed((12345,) * 6)
    G.steps(100 * cfg.system.P)
    rows = []
    for b in range(nblocks):
        G.accum_reset()
        for k in range(per_block // skip):
            G.steps(skip, sync=False)
            G.measure()
        G.sync()
        s = G.block_scalars()
        rows.append([s.kin / s.count, s.pot / s.count, s.rot / s.count])
    G.close()
    return np.array(rows)


def compare(g, r, cols):
    for i, nm in cols:
        mg, mr = g[:, i].mean(), r[:, i].mean()
        sg, sr = g[:, i].std(ddof=1) / np.sqrt(len(g)), r[:, i].std(ddof=1) / np.sqrt(len(r))
        sigma = np.hypot(sg, sr)
        print(f"{nm}: gpu {mg:.5f} +- {sg:.5f}   reference {mr:.5f} +- {sr:.5f}   diff {abs(mg-mr)/sigma:.2f} sigma")
        assert abs(mg - mr) < 2.0 * sigma + 1e-9 * abs(mr), f"{nm}: {mg} vs {mr} (sigma {sigma})"


def test_free_linear_rotor_matches_reference(pkg):
    """examples/CO2_100K_4_4: zero potential, <E_rot> is set by the tabulated density matrix alone."""
    kw = {}
    r = reference_blocks("CO2", kw, nblocks=20, per_block=20000, skip=4)
    g = gpu_blocks(pkg, "CO2", kw, nchains=32, nblocks=20, per_block=2000, skip=4)
    compare(g, r, [(0, "K"), (2, "E_rot")])


def test_rotor_in_solvent_cluster_matches_reference(pkg):
    """Reduced C5 (N2O + 6 pH2, P=64, Q=16, 2 K: a bound cluster with 28 % / 69 % / 53 % bisection / bisection / rotation
    acceptance): all three move types and both tabulated potentials."""
    kw = dict(P=64, Q=16, nsolv=6, temperature=2.0)
    r = reference_blocks("C5", kw, nblocks=20, per_block=12800, skip=16)
    g = gpu_blocks(pkg, "C5", kw, nchains=32, nblocks=20, per_block=3200, skip=16)
    compare(g, r, [(0, "K"), (1, "V"), (2, "E_rot")])


def test_cxx_driver_writes_reference_formats(pkg, tmp_path):
    """pimc_b200 on the reference's CO2 deck: .eng rows in the reference's column layout (mc_main.cc:780-792)."""
    drv = os.path.join(ROOT, "moribs-pimc_b200", "driver", "pimc_b200")
    if not os.path.exists(drv):
        pytest.skip("driver binary not built")
    d = os.path.join(pkg.configs.DECKS, "CO2_100K_4_4")
    for f in ("CO2_T100t4.rot", "CO2_fake.pot"):
        shutil.copy(os.path.join(d, f), tmp_path)
    deck = open(os.path.join(d, "qmc.input")).read().replace("NUMBEROFBLOCKS     2000  500", "NUMBEROFBLOCKS     6  2")
    deck = deck.replace("OUTPUTDIR        ./g4/1/", "OUTPUTDIR        ./")
    open(tmp_path / "qmc.input", "w").write(deck)
    out = subprocess.run([drv, "--chains", "16"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    rows = [l for l in open(tmp_path / "CO2_monomer.eng")]
    assert len(rows) == 4                                   # blocks 3..6 (two equilibration blocks)
    first = rows[0]
    assert first[:4].strip() == "3" and len(first.split()) == 10
    assert all(len(x) in (12, 13) and "e" in x for x in first.split()[1:])      # setprecision(6), scientific
    vals = np.array([[float(x) for x in l.split()] for l in rows])
    assert np.all(np.abs(vals[:, 1] - 150.0) < 5.0)         # kinetic energy of one free particle at 100 K
    assert np.all(np.abs(vals[:, 4] - 97.0) < 6.0)          # rotational energy of CO2 at 100 K, 4 slices
    assert np.allclose(vals[:, 3], vals[:, 1] + vals[:, 2]) and np.allclose(vals[:, 6], vals[:, 3] + vals[:, 4], rtol=1e-5)
    for f in ("yw001.stat", "yw001.conf", "yw001.tabl", "CO2_monomer.xyz", "CO2_monomer_sum.eng", "CO2_monomer_sum.rcf", "CO2_monomer003.rcf"):
        assert os.path.exists(tmp_path / f), f
    assert open(tmp_path / "yw001.stat").read().startswith("STARTBLOCK ")
    assert os.path.getsize(tmp_path / "yw001.conf") == 8 + 2 * 8 * 4      # streamsize + x row + cosine x row (N*P = 4)
