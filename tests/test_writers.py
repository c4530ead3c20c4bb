"""CPU test of the driver's density writers (moribs-pimc_b200/driver/pimc_writers.h) against the reference's OWN
Save* functions (SaveDensities1D, SaveDensities2D, SaveRho1D, SaveRhoThetaChi, SaveDensities3D; mc_estim.cc:1327-1820,
1930-1995; SaveEnergy, SaveSumEnergy mc_main.cc:764-836; SaveRCF, SaveGraSum, SaveExchangeLength, SaveAreaEstimators,
SaveAreaEstim3D mc_estim.cc:1141-1191,1288-1326,2021-2085,2596-2729) driven through oracle/_ref on the same seeded
histograms and accumulator values: every output file must be byte-identical.
Where oracle/_ref is absent the files are checked against the digests committed in tests/golden/writers.json
(made by this test when run with MAKE_WRITER_FIXTURE=1 next to the reference)."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "writers.json")
CASES = {"C5": dict(P=32, Q=8, nsolv=6), "C1": dict(P=64, Q=16), "C2": dict(P=32, Q=8, nsolv=5), "C4": dict(P=64, Q=32)}

SCRIPT = r'''
import sys, os, json, hashlib, ctypes as C, subprocess
import numpy as np
sys.path.insert(0, %(root)r)
name, work, use_ref = %(name)r, %(work)r, %(use_ref)r
import __graft_entry__ as ge
pkg = ge.load_package()
cfg = pkg.configs.make_config(name, big_tables=(name not in ("C1", "C2", "C4")), **%(kw)r)
if name in ("C1", "C2"):                                  # tables are irrelevant for the writers; keep the process small
    cfg.tables["pot3d"] = (3, 181, 181, 4.0, 20.0, np.zeros(3 * 181 * 181))
if name in ("C1", "C2", "C4"):
    z = np.zeros(181 * 361 * 361)
    cfg.tables["rot3d"] = (z, z, z)
s = cfg.system
rng = np.random.default_rng(11)
g1 = rng.integers(0, 500, 300).astype(float)
g2 = rng.integers(0, 60, 300 * 50).astype(float)
g3 = rng.integers(0, 7, 300 * 50 * 100).astype(float)
rel = rng.integers(0, 900, 250).astype(float)
acount = 1234.0
dp = C.POINTER(C.c_double)
P_ = lambda a: a.ctypes.data_as(dp)
os.makedirs(work + "/mine", exist_ok=True); os.makedirs(work + "/ref", exist_ok=True)
shim = C.CDLL(%(shim)r)
numb = np.array([t.numb for t in s.types], dtype=np.int32); mol = np.array([t.molecule for t in s.types], dtype=np.int32)
box = (s.N / s.density) ** (1.0 / 3.0)
shim.shim_save_densities((work + "/mine/gr").encode(), C.c_int(s.P), C.c_int(s.Q), C.c_int(len(s.types)), numb.ctypes.data_as(C.POINTER(C.c_int)),
                         mol.ctypes.data_as(C.POINTER(C.c_int)), C.c_double(box ** 3), C.c_double(acount), P_(g1), P_(g2), P_(g3), P_(rel))
# per-block scalar writers: SaveEnergy, SaveSumEnergy, SaveRCF (block + total), SaveGraSum, SaveExchangeLength, SaveAreaEstimators, SaveAreaEstim3D
scal7 = np.array([123.456, -987.654, 3.25, 17.5, -4321.0, -2222.0, 11.0]) * acount
Qn = max(1, s.Q)
rcf0 = rng.uniform(-1, 1, Qn) * acount * Qn
rcf19 = rng.integers(0, Qn, Qn).astype(float) * acount
bos = [t for t in s.types if t.stat == 1]
nb = bos[0].numb if bos else 0
ploops = rng.integers(0, 50, max(1, nb)).astype(float)
pindex = rng.permutation(max(1, nb)).astype(np.int32)
area40 = rng.uniform(0.5, 3.0, 40) * acount
imol = [t for t in s.types if t.molecule]
linear = int(bool(imol) and imol[0].molecule == 1); mff = int(bool(imol) and imol[0].molecule == 2 and not s.ispher)
atoms = [t for t in s.types if not t.molecule]
bmass = bos[0].mass if bos else 1.0
lam = 0.5 * (100.0 * (1.05457266 * 1.05457266) / (1.6605402 * 1.380658)) / bmass            # mc_setup.cc:206-215
IP_ = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
shim.shim_save_block((work + "/mine/gr").encode(), C.c_long(17), C.c_int(s.N), C.c_int(s.P), C.c_int(s.Q), C.c_double(s.temperature),
                     C.c_int(1 if atoms else 0), C.c_int(atoms[0].numb if atoms else 0), C.c_int(nb), C.c_int(linear), C.c_int(mff),
                     C.c_double(lam), C.c_double(bmass), C.c_double(acount), P_(scal7), P_(rcf0), P_(rcf19), P_(g1), P_(ploops), IP_(pindex), P_(area40))
# IOxyz / IOxyzAng of one configuration (cosine = the unit axis the reference derives from the angles)
cs = np.zeros_like(cfg.angles)
st = np.sqrt(np.maximum(0.0, 1.0 - cfg.angles[1] ** 2))
cs[0], cs[1], cs[2] = st * np.cos(cfg.angles[0]), st * np.sin(cfg.angles[0]), cfg.angles[1]
perm_all = np.arange(s.N, dtype=np.int32)
if nb:
    perm_all[:nb] = pindex
names = (C.c_char_p * len(s.types))(*[t.name.encode() for t in s.types])
coords_c = np.ascontiguousarray(cfg.coords); angles_c = np.ascontiguousarray(cfg.angles)
out = {}
if use_ref:
    from oracle import oracle_py as op
    R = op.Ref(cfg)
    R.lib.ref_save_densities((work + "/ref/gr").encode(), C.c_double(acount), P_(g1), P_(g2), P_(g3), P_(rel))
    if bos:
        assert abs(R.lib.ref_lambda(s.types.index(bos[0])) - lam) < 1e-12 * lam
    R.set_state(coords_c, angles_c, perm_all)
    co, ao, cs_ref = R.get_state()                    # the reference's own MCCosine for the IOxyz columns
    R.lib.ref_write_xyz((work + "/ref/gr.xyz").encode(), (work + "/ref/gr017").encode())
    shim.shim_write_xyz((work + "/mine/gr.xyz").encode(), (work + "/mine/gr017").encode(), C.c_int(len(s.types)), names, IP_(numb), C.c_int(s.P),
                        P_(coords_c), P_(angles_c), P_(np.ascontiguousarray(cs_ref)), C.c_int(nb), IP_(perm_all))
    R.lib.ref_save_block((work + "/ref/gr").encode(), C.c_long(17), C.c_double(acount), P_(scal7), P_(rcf0), P_(rcf19), P_(g1), P_(ploops), IP_(pindex), P_(area40))
    for f in sorted(os.listdir(work + "/ref")):
        a = open(work + "/ref/" + f, "rb").read()
        b = open(work + "/mine/" + f, "rb").read() if os.path.exists(work + "/mine/" + f) else b""
        out[f] = {"identical": a == b, "bytes": len(a), "md5": hashlib.md5(a).hexdigest()}
    out["_only_mine"] = sorted(set(os.listdir(work + "/mine")) - set(os.listdir(work + "/ref")))
else:
    shim.shim_write_xyz((work + "/mine/gr.xyz").encode(), (work + "/mine/gr017").encode(), C.c_int(len(s.types)), names, IP_(numb), C.c_int(s.P),
                        P_(coords_c), P_(angles_c), P_(np.ascontiguousarray(cs)), C.c_int(nb), IP_(perm_all))
    for f in sorted(os.listdir(work + "/mine")):
        a = open(work + "/mine/" + f, "rb").read()
        out[f] = {"bytes": len(a), "md5": hashlib.md5(a).hexdigest()}
print("RESULT " + json.dumps(out))
'''


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "libwshim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "writers_shim.cpp")])
    return so


@pytest.mark.parametrize("name", list(CASES))
def test_density_writers_byte_identical_to_reference(name, shim, tmp_path):
    use_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpimcref.so"))
    code = SCRIPT % dict(root=ROOT, name=name, work=str(tmp_path), use_ref=use_ref, kw=CASES[name], shim=shim)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(line[-1][7:])
    top = {"gr.gra", "gr.gri", "gr.grt", "gr.grc", "gr.gtc", "gr_sum.g3d", "gr_sum.gri", "gr_sum.grt", "gr_sum.grc", "gr_sum.eulphi", "gr_sum.eulchi", "gr_sum.eulthe"}
    common = {"gr.eng", "gr_sum.eng", "gr.rcf", "gr_sum.rcf", "gr_sum.gra", "gr.xyz", "gr017.xyz"}
    expect = {"C5": {"gr.gra", "gr.gri", "gr.grt", "gr.g2d", "gr_sum.g2d", "gr.prl", "gr.sup", "gr.sffs3d"} | common,
              "C1": top | common | {"gr.prl", "gr.sffs3d", "gr.mffs3d"}, "C2": top | common | {"gr.prl", "gr.sffs3d", "gr.mffs3d"}, "C4": top | common}
    fixture = json.load(open(FIXTURE)) if os.path.exists(FIXTURE) else {}
    if use_ref:
        assert res.pop("_only_mine") == []
        assert set(res) == expect[name]
        for f, r in res.items():
            assert r["identical"], f"{name}: {f} differs from the reference's writer"
            assert r["bytes"] > 50
        if os.environ.get("MAKE_WRITER_FIXTURE"):
            fixture[name] = {f: {"bytes": r["bytes"], "md5": r["md5"]} for f, r in res.items()}
            json.dump(fixture, open(FIXTURE, "w"), indent=1, sort_keys=True)
        elif name in fixture:
            assert {f: r["md5"] for f, r in res.items()} == {f: r["md5"] for f, r in fixture[name].items()}
    else:
        if name not in fixture:
            pytest.skip("neither oracle/_ref nor the digest fixture is available")
        assert set(res) == set(fixture[name])
        for f, r in res.items():
            assert r["md5"] == fixture[name][f]["md5"], f"{name}: {f} differs from the reference's writer (digest)"
