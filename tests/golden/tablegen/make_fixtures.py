"""Fixtures for the rho-table generators, made from the reference's OWN golden outputs.

Run here (needs /root/reference):  python tests/golden/tablegen/make_fixtures.py

  ref_asymrho_den010.npz   nmv_prop/rho.den010_{rho,eng,esq}: asymrho.x 0.37 128 -1 10 10 0.6666525 0.2306476 0.1769383 66
                           (nmv_prop/a-run:4) + the partition-function lines of nmv_prop/log
  ref_symrho_den0{00,10}.npz  symtop_prop/rho.den0{00,10}_*: symrho.x 0.37 128 1 <ith> <ith> 0.5 0.3 66 (symtop_prop/a-run:4)
  ref_linden_N2O.npz / ref_linden_CO2.npz  examples/*/N2O_T0.5t128.rot, CO2_T100t4.rot (linden.out of linden.f)
  oracle_asym_small.npz    oracle/libtablegen_oracle.so outputs for a small asymmetric top (whole plane incl. symmetry fill,
                           iodevn -1/0/1) and Wigner-d matrices in quad precision: travels to the GPU box as the tight
                           reference of the device generators

The files hold the text values parsed to float64; E15.8 carries 8 significant digits, so formatting a parsed value
reproduces the reference's text byte for byte (checked below).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", "..", ".."))
sys.path.insert(0, ROOT)
from oracle import tablegen_py as tg  # noqa: E402

REF = "/root/reference"


def plane(path):
    txt = open(path).read().split("\n")[:-1]
    v = np.array([float(t) for t in txt])
    assert v.size == 361 * 361
    assert all(tg.fmt_e15_8(x) == t for x, t in zip(v[::97], txt[::97])), path
    return v.reshape(361, 361)


def main():
    log = open(f"{REF}/nmv_prop/log").read().split("\n")
    at = {}
    for i, l in enumerate(log):
        if l.startswith("AT BETA") or l.startswith("AT TAU"):
            rows = []
            for r in log[i + 1:i + 4]:
                toks = r.replace("=", " ").split()
                rows.append([float(t) for t in toks if t.replace(".", "").replace("-", "").isdigit()])
            at[l.strip()] = rows
    np.savez_compressed(f"{HERE}/ref_asymrho_den010.npz",
                        args=np.array([0.37, 128, -1, 10, 10, 0.6666525, 0.2306476, 0.1769383, 66]),
                        rho=plane(f"{REF}/nmv_prop/rho.den010_rho"), eng=plane(f"{REF}/nmv_prop/rho.den010_eng"),
                        esq=plane(f"{REF}/nmv_prop/rho.den010_esq"),
                        at_beta=np.array([r[:4] for r in at["AT BETA"]]),     # Z, E (cm-1), E (K), Cv for even k, odd k, classical
                        at_tau=np.array([r[:3] for r in at["AT TAU"]]))
    for ith in (0, 10):
        d = {"args": np.array([0.37, 128, 1, ith, ith, 0.5, 0.3, 66])}
        for nm in ("rho", "eng", "esq"):
            p = f"{REF}/symtop_prop/rho.den{ith:03d}_{nm}"
            if os.path.exists(p):
                d[nm] = plane(p)
        np.savez_compressed(f"{HERE}/ref_symrho_den{ith:03d}.npz", **d)
    for name, path, args in (("N2O", f"{REF}/examples/N2O_5pH2_0.5K_512_128/N2O_T0.5t128.rot", [0.5, 128, 0.419, 1400, -1]),
                             ("CO2", f"{REF}/examples/CO2_100K_4_4/CO2_T100t4.rot", [100.0, 4, 0.39021, 3000, -1])):
        lines = [l for l in open(path) if not l.startswith("#")]
        v = np.loadtxt(path)
        assert tg.rot_lines(v) == lines, path          # text <-> float64 round trip is the identity
        np.savez_compressed(f"{HERE}/ref_linden_{name}.npz", args=np.array(args), table=v)
    # small asymmetric top from the oracle (quad-precision Wigner d), whole planes
    small = dict(T=10.0, nslice=2, A=27.877, B=14.512, C=9.285, maxj=10)
    d = {"args": np.array([small["T"], small["nslice"], small["A"], small["B"], small["C"], small["maxj"]])}
    for io in (-1, 0, 1):
        o = tg.AsymRho(small["T"], small["nslice"], io, small["A"], small["B"], small["C"], small["maxj"])
        for ith in (0, 37, 90, 180):
            r, e, q = o.plane(ith)
            d[f"io{io}_th{ith}"] = np.stack([r, e, q])[:, ::5, ::5]      # every 5th degree of the filled plane
        d[f"io{io}_info"] = o.info
        d["eng_even"], d["eng_odd"] = np.sort(o.energies(0)), np.sort(o.energies(1))
        o.close()
    for maxj, ths, js in ((24, (0.3, np.pi / 2, 2.5), range(25)), (66, (10 * np.pi / 180, 3.0), (30, 50, 66))):
        for it, th in enumerate(ths):
            w = 2 * maxj + 1
            dm = np.zeros((len(js), w, w))
            for ij, j in enumerate(js):
                for m in range(-j, j + 1):
                    for k in range(-j, j + 1):
                        dm[ij, m + maxj, k + maxj] = tg.wigd(j, m, k, th)
            d[f"wigd_j{maxj}_{it}"] = dm
            d[f"wigd_j{maxj}_{it}_theta"] = np.array(th)
            d[f"wigd_j{maxj}_{it}_js"] = np.array(list(js))
    np.savez_compressed(f"{HERE}/oracle_asym_small.npz", **d)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
