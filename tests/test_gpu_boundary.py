"""Reference-side proof of the drop-in boundary: the REFERENCE's own mc_main.cc -- its input
parser, set-up, block loop, Save* writers and checkpoint code -- with the patch of INTEGRATION.md section B applied to a
build-time copy (oracle/make_patched_main.py -> oracle/_ref/pimc_ref_gpu, linked against libpimcgpu.so) runs the
reference's CPU-runnable example deck examples/CO2_100K_4_4 next to this repo's pimc_b200.  Same start configuration
(xyz.init through the reference's initconf_), same MRG32k3a package seed, one chain: both programs drive the same
library through the same C ABI, so their .eng rows must agree to the printed digits."""
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _deck(pkg, tmp, blocks="6  2", passes="2000"):
    d = os.path.join(pkg.configs.DECKS, "CO2_100K_4_4")
    for f in ("CO2_T100t4.rot", "CO2_fake.pot"):
        shutil.copy(os.path.join(d, f), tmp)
    deck = open(os.path.join(d, "qmc.input")).read()
    deck = deck.replace("NUMBEROFBLOCKS     2000  500", f"NUMBEROFBLOCKS     {blocks}").replace("NUMBEROFPASSES     3000 ", f"NUMBEROFPASSES     {passes} ")
    deck = deck.replace("OUTPUTDIR        ./g4/1/", "OUTPUTDIR        ./")
    assert f"NUMBEROFBLOCKS     {blocks}" in deck and f"NUMBEROFPASSES     {passes}" in deck
    open(os.path.join(tmp, "qmc.input"), "w").write(deck + "READMCCOORDS\n")
    rng = np.random.default_rng(4)
    with open(os.path.join(tmp, "xyz.init"), "w") as f:            # initconf.f:10-23: count (+ permutation), comment, label x phi y cos(theta) z chi
        f.write("4\n# start configuration of the boundary test\n")
        phi0, ct0 = rng.uniform(0, 6.28), rng.uniform(-0.8, 0.8)         # one orientation, slightly different on the four slices
        for it in range(4):
            x, y, z = 0.05 * rng.standard_normal(3)
            f.write(f"C {x:.6e} {phi0 + 0.05 * rng.standard_normal():.6e} {y:.6e} {ct0 + 0.02 * rng.standard_normal():.6e} {z:.6e} 0.000000e+00\n")


def test_reference_main_through_the_c_abi_matches_pimc_b200(pkg, tmp_path):
    ref_gpu = os.path.join(ROOT, "oracle", "_ref", "pimc_ref_gpu")
    drv = os.path.join(ROOT, "moribs-pimc_b200", "driver", "pimc_b200")
    if not os.path.exists(ref_gpu):
        pytest.skip("oracle/_ref/pimc_ref_gpu not built (make -C oracle ref_gpu needs /root/reference at build time)")
    if not os.path.exists(drv):
        pytest.skip("driver binary not built")
    a, b = tmp_path / "reference_main", tmp_path / "pimc_b200"
    a.mkdir(); b.mkdir()
    _deck(pkg, a); _deck(pkg, b)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    ra = subprocess.run([ref_gpu], cwd=a, env=env, capture_output=True, text=True, timeout=900)
    assert ra.returncode in (0, 1), ra.stdout[-3000:] + ra.stderr[-2000:]            # the reference's main returns 1 on success
    rb = subprocess.run([drv, "--chains", "1"], cwd=b, env=env, capture_output=True, text=True, timeout=900)
    assert rb.returncode == 0, rb.stdout[-3000:] + rb.stderr[-2000:]
    ea = np.loadtxt(a / "CO2_monomer.eng", ndmin=2)
    eb = np.loadtxt(b / "CO2_monomer.eng", ndmin=2)
    print("reference main through libpimcgpu.so:\n", ea, "\npimc_b200:\n", eb)
    assert ea.shape == eb.shape == (4, 10) and list(ea[:, 0]) == [3, 4, 5, 6]
    # columns written by the reference's own SaveEnergy from the accumulators fetched over the ABI (mc_main.cc:780-792)
    assert np.allclose(ea[:, 1:8], eb[:, 1:8], rtol=2e-6, atol=1e-9), np.abs(ea - eb).max()
    assert abs(ea[:, 1].mean() - 150.0) < 30.0 and abs(ea[:, 4].mean() - 97.0) < 30.0            # free particle at 100 K; CO2 rotor, 4 slices (8000 steps per block)
    # the reference's checkpoint and configuration writers ran on the state downloaded through the ABI
    for f in ("yw001.stat", "yw001.conf", "yw001.tabl", "CO2_monomer.xyz", "CO2_monomer003.rcf"):
        assert os.path.exists(a / f), f
    ra_rcf = np.loadtxt(a / "CO2_monomer006.rcf", max_rows=4)
    rb_rcf = np.loadtxt(b / "CO2_monomer006.rcf", max_rows=4)
    assert np.allclose(ra_rcf[:, :2], rb_rcf[:, :2], rtol=2e-6)
