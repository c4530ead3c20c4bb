"""System descriptions, input decks and synthetic tables for the PIMC hot path.

Host-side mirror of the reference's input layer for the five BASELINE
configurations: the ``qmc.input`` keyword deck
(mc_input.cc:18-57,115-343), the 1-D potential / linear-rotor density file
readers (mc_poten.cc:757-865), ``xyz.init`` (initconf.f:1-27) and the
synthetic stand-ins for the tables that are git-LFS pointers in the reference
(SURVEY.md section 8d).  Pure numpy; no GPU code here.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import Dict, List, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DECKS = os.path.join(HERE, "data", "decks")          # the reference's example inputs (qmc.input, *.pot, *.rot, xyz.init) used by the harness

# mc_const.h:50-56 and mc_setup.cc:244-319 (MCInitParams)
_H1, _H2, _HE4, _C12, _N14, _O16, _S32 = 1.0078, 2.015650642, 4.0026032497, 12.0, 14.003, 15.994915, 31.972
MASS: Dict[str, float] = {
    "He4": _HE4, "H2": _H2, "OCS": _O16 + _C12 + _S32, "N2O": 2.0 * _N14 + _O16, "CO2": _C12 + 2.0 * _O16,
    "CO": _C12 + _O16, "HCN": _H1 + _C12 + _N14, "HCCCN": _H1 + 3.0 * _C12 + _N14, "H2O": 2.0 * _H1 + _O16,
    "SO2": 2.0 * _O16 + _S32, "HCOOCH3": 4.0 * _H1 + 2.0 * _O16 + 2.0 * _C12,
}
WNO2K = 0.6950356
SIZE_ROTDEN = 181 * 361 * 361


@dataclasses.dataclass
class PType:
    name: str
    numb: int
    molecule: int          # 0 atom, 1 linear, 2 non-linear
    stat: int              # 0 BOLTZMANN, 1 BOSE
    mcstep: float
    levels: int
    fpot: str
    rtstep: float = 0.0

    @property
    def mass(self) -> float:
        return MASS[self.name]


@dataclasses.dataclass
class System:
    types: List[PType]
    P: int
    Q: int
    temperature: float
    density: float = 0.02
    ispher: int = 0
    minimage: int = 0
    rotden_type: int = 0
    rot_odevn: int = 0
    rot_eoff: float = 0.0
    x_rot: float = 0.0
    y_rot: float = 0.0
    z_rot: float = 0.0
    rnratio: int = 1
    worm: Optional[tuple] = None
    read_coords: bool = False
    reflect: tuple = (0, 0, 0)       # IREFLX, IREFLY, IREFLZ
    rotsym: int = 0
    passes: int = 1
    blocks: int = 1
    eq_blocks: int = 0
    skip_averg: int = 1
    prefix: str = "pimc"

    @property
    def N(self) -> int:
        return sum(t.numb for t in self.types)

    @property
    def R(self) -> int:
        return self.P // self.Q if self.Q else 1

    @property
    def tau(self) -> float:
        return 1.0 / self.temperature / self.P

    def bead_updates_per_pass(self) -> Dict[str, int]:
        """bisection 2^L-1 per atom per call, molecular P per atom, rotation 1 per step."""
        bis = sum(self.P * t.numb * ((1 << t.levels) - 1) for t in self.types)
        mol = sum(self.P * t.numb for t in self.types)
        rot = sum(self.P * self.Q * t.numb for t in self.types if t.molecule and self.Q)
        return {"bisection": bis, "molecular": mol, "rotation": rot, "total": bis + mol + rot}


def parse_qmc_input(path: str) -> System:
    """Keyword deck reader with the reference's semantics (mc_input.cc:115-343)."""
    types: List[PType] = []
    kw: Dict[str, List[str]] = {}
    rot = None
    for raw in open(path):
        tok = raw.split("#", 1)[0].split() if not raw.lstrip().startswith("#") else []
        if not tok:
            continue
        key = tok[0]
        if key in ("ATOM", "MOLECULE", "NONLINEAR"):
            numb = int(tok[2])
            if numb < 0:
                kw["ISPHER"] = ["1"]
                numb = -numb
            stat = {"BOLTZMANN": 0, "BOSE": 1}[tok[3]]
            mol = {"ATOM": 0, "MOLECULE": 1, "NONLINEAR": 2}[key]
            if numb > 0:
                types.append(PType(tok[1], numb, mol, stat, float(tok[4]), int(tok[5]), tok[6]))
        elif key == "ROTATION":
            rot = (tok[1], float(tok[2]), int(tok[3]))
        else:
            kw[key] = tok[1:]
    Q = 0
    if rot:
        for t in types:
            if t.name == rot[0]:
                t.rtstep = rot[1]
        Q = rot[2]
    s = System(types=types, P=int(kw["NUMBEROFSLICES"][0]), Q=Q, temperature=float(kw["TEMPERATURE"][0]),
               density=float(kw.get("DENSITY", ["0.02"])[0]))
    s.ispher = int(kw.get("ISPHER", ["0"])[0])
    s.minimage = 1 if "MINIMAGE" in kw else 0
    if "ROTDENSI" in kw:
        r = kw["ROTDENSI"]
        s.rotden_type, s.rot_odevn, s.rot_eoff = int(r[0]), int(r[1]), float(r[2])
        s.x_rot, s.y_rot, s.z_rot, s.rnratio = float(r[3]), float(r[4]), float(r[5]), int(r[6])
    if "WORM" in kw:
        s.worm = (kw["WORM"][0], float(kw["WORM"][1]), int(kw["WORM"][2]))
    s.read_coords = "READMCCOORDS" in kw
    s.reflect = tuple(int(kw.get(k, ["0"])[0]) for k in ("REFLECTX", "REFLECTY", "REFLECTZ"))
    s.rotsym = int(kw["ROTSYM"][0]) if "ROTSYM" in kw else 0
    s.passes = int(kw.get("NUMBEROFPASSES", ["1"])[0])
    if "NUMBEROFBLOCKS" in kw:
        s.blocks, s.eq_blocks = int(kw["NUMBEROFBLOCKS"][0]), int(kw["NUMBEROFBLOCKS"][1])
    s.skip_averg = int(kw.get("MCSKIP_AVERG", ["1"])[0])
    s.prefix = kw.get("FILENAMEPREFIX", ["pimc"])[0]
    return s


def write_qmc_input(s: System, path: str, outdir: str = "./out/") -> None:
    """Emit a deck the reference's own parser accepts (used to drive oracle/_ref)."""
    lines = ["MASTERDIR ./", f"OUTPUTDIR {outdir}", f"FILENAMEPREFIX {s.prefix}", "DIMENSION 3",
             f"DENSITY {s.density!r}", f"TEMPERATURE {s.temperature!r}"]
    for t in s.types:
        key = {0: "ATOM", 1: "MOLECULE", 2: "NONLINEAR"}[t.molecule]
        numb = -t.numb if (t.molecule == 2 and s.ispher) else t.numb
        lines.append(f"{key} {t.name} {numb} {'BOSE' if t.stat else 'BOLTZMANN'} {t.mcstep!r} {t.levels} {t.fpot} PRIMITIVE")
    for t in s.types:
        if t.molecule and s.Q:
            lines.append(f"ROTATION {t.name} {t.rtstep!r} {s.Q}")
    if s.rotden_type or s.x_rot:
        lines.append(f"ROTDENSI {s.rotden_type} {s.rot_odevn} {s.rot_eoff!r} {s.x_rot!r} {s.y_rot!r} {s.z_rot!r} {s.rnratio}")
    if s.worm:
        lines.append(f"WORM {s.worm[0]} {s.worm[1]!r} {s.worm[2]}")
    if s.minimage:
        lines.append("MINIMAGE")
    for key, flag in zip(("REFLECTX", "REFLECTY", "REFLECTZ"), s.reflect):      # symmetry moves of MCGetAverage, mc_main.cc:647-692
        if flag:
            lines.append(f"{key} {flag}")
    if s.rotsym:
        lines.append(f"ROTSYM {s.rotsym}")
    lines += [f"NUMBEROFSLICES {s.P}", f"NUMBEROFPASSES {s.passes}", f"NUMBEROFBLOCKS {s.blocks} {s.eq_blocks}",
              "MCSKIP_RATIO 100000000", "MCSKIP_TOTAL 100000000", f"MCSKIP_AVERG {s.skip_averg}"]
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


# ----------------------------------------------------------------------------
# table file readers
# ----------------------------------------------------------------------------
def load_columns(path: str, ncol: int) -> np.ndarray:
    """read_datafile (mc_poten.cc:757-865): first `ncol` tokens of every line whose first token is not '#'."""
    rows = []
    for line in open(path):
        tok = line.split()
        if not tok or tok[0] == "#":
            continue
        rows.append([float(x) for x in tok[:ncol]])
    return np.ascontiguousarray(np.array(rows, dtype=np.float64).T)


def load_vspher_table() -> np.ndarray:
    """the 501-entry DATA table of vspher.f:15-517 as the driver compiles it in (data/vspher_table.h, hex doubles)"""
    import re
    txt = open(os.path.join(HERE, "data", "vspher_table.h")).read()
    v = np.array([float.fromhex(x) for x in re.findall(r"-?0x[0-9a-f.]+p[-+]\d+", txt)])
    assert len(v) == 501
    return v


def load_xyz_init(path: str, nbeads: int, nboson: int):
    """initconf.f:1-27 + mc_main.cc:184-201 -> coords[3][N*P], angles[3][N*P] (phi, cos(theta), chi), perm."""
    with open(path) as f:
        head = f.readline().split()
        perm = np.array([int(x) for x in head[1:1 + nboson]], dtype=np.int32)
        f.readline()
        data = np.loadtxt(f, usecols=(1, 2, 3, 4, 5, 6), max_rows=nbeads)
    coords = np.ascontiguousarray(data[:, 0::2].T)
    angles = np.ascontiguousarray(data[:, 1::2].T)
    return coords, angles, perm


# ----------------------------------------------------------------------------
# synthetic tables
# ----------------------------------------------------------------------------
def synth_rot3d(temperature: float, Q: int, A: float, B: float, C: float):
    """rho = exp(-k(3-trR)), erot = a + b(3-trR) [cm^-1], esq = erot^2; theta outer, phi, chi inner, 1-degree grid."""
    tau = 1.0 / (WNO2K * temperature * Q)
    bbar = (A + B + C) / 3.0
    kappa = 1.0 / (4.0 * bbar * tau)
    a = 1.5 / tau
    b = -1.0 / (4.0 * bbar * tau * tau)
    th = np.deg2rad(np.arange(181.0))[:, None, None]
    ph = np.deg2rad(np.arange(361.0))[None, :, None]
    ch = np.deg2rad(np.arange(361.0))[None, None, :]
    x = 3.0 - ((1.0 + np.cos(th)) * (1.0 + np.cos(ph + ch)) - 1.0)
    rho = np.exp(-kappa * x).reshape(-1)
    erot = (a + b * x).reshape(-1)
    esq = erot * erot
    return np.ascontiguousarray(rho), np.ascontiguousarray(erot), np.ascontiguousarray(esq)


def synth_pot3d(rg: int = 501, thg: int = 181, chg: int = 181, rmin: float = 4.0, rmax: float = 20.0,
                eps: float = 30.0, sigma0_bohr: float = 6.614, a: float = 0.15, b: float = 0.05) -> np.ndarray:
    """V(r,theta,chi) = 4 eps [(s/r)^12 - (s/r)^6], s = s0 (1 + a cos^2 th + b sin^2 th cos 2chi); r in bohr."""
    r = np.linspace(rmin, rmax, rg)[:, None, None]
    th = np.deg2rad(np.arange(float(thg)))[None, :, None]
    ch = np.deg2rad(np.arange(float(chg)))[None, None, :]
    s = sigma0_bohr * (1.0 + a * np.cos(th) ** 2 + b * np.sin(th) ** 2 * np.cos(2.0 * ch))
    x6 = (s / r) ** 6
    return np.ascontiguousarray((4.0 * eps * (x6 * x6 - x6)).reshape(-1))


def synth_pot2d(rsize: int = 2001, csize: int = 1001, dr: float = 0.005, dc: float = 0.002, r0: float = 2.0,
                eps: float = 30.0, sigma0: float = 3.5, a: float = 0.15):
    rgrid = r0 + dr * np.arange(rsize)
    cgrid = -1.0 + dc * np.arange(csize)
    s = sigma0 * (1.0 + a * cgrid[None, :] ** 2)
    x6 = (s / rgrid[:, None]) ** 6
    return rgrid, cgrid, np.ascontiguousarray(4.0 * eps * (x6 * x6 - x6))


def write_pot2d(path: str, rgrid, cgrid, v, dr, dc) -> None:
    """2-D potential file in the README.md:78-123 format read by init_pot2D (mc_poten.cc:338-360)."""
    with open(path, "w") as f:
        f.write(f"{len(rgrid)} {len(cgrid)}\n{dr!r} {dc!r}\n")
        f.write(" ".join(repr(float(x)) for x in rgrid) + "\n")
        f.write(" ".join(repr(float(x)) for x in cgrid) + "\n")
        for row in v:
            f.write(" ".join(repr(float(x)) for x in row) + "\n")


# ----------------------------------------------------------------------------
# initial configurations
# ----------------------------------------------------------------------------
def cluster_config(s: System, seed: int = 1, spacing: float = 3.8, core: float = 3.3, jitter: float = 0.05):
    """Classical start: rotor(s) near the origin, solvent atoms on the nearest sites of a cubic lattice.

    All P beads of a particle start at the same point plus a small Gaussian jitter.  Angles: phi, chi
    uniform, cos(theta) uniform (identical on all rot slices + jitter)."""
    rng = np.random.default_rng(seed)
    N, P = s.N, s.P
    coords = np.zeros((3, N * P))
    angles = np.zeros((3, N * P))
    angles[1, :] = 1.0
    nsolv = sum(t.numb for t in s.types if t.molecule == 0)
    nmol = N - nsolv
    m = 8
    g = np.arange(-m, m + 1) * spacing
    sites = np.array(np.meshgrid(g, g, g, indexing="ij")).reshape(3, -1).T
    d = np.linalg.norm(sites, axis=1)
    sites = sites[d > core]
    sites = sites[np.argsort(np.linalg.norm(sites, axis=1), kind="stable")]
    atom = 0
    isolv = 0
    for t in s.types:
        for k in range(t.numb):
            if t.molecule == 0:
                c = sites[isolv]
                isolv += 1
            else:
                c = np.array([2.9 * (k - 0.5 * (nmol - 1)), 0.0, 0.0])
            sl = slice(atom * P, (atom + 1) * P)
            coords[:, sl] = c[:, None] + jitter * rng.standard_normal((3, P))
            if t.molecule:
                Q = s.Q
                phi0, ct0, chi0 = rng.uniform(0, 2 * math.pi), rng.uniform(-0.9, 0.9), rng.uniform(0, 2 * math.pi)
                q = slice(atom * P, atom * P + Q)
                angles[0, q] = np.mod(phi0 + 0.05 * rng.standard_normal(Q), 2 * math.pi)
                angles[1, q] = np.clip(ct0 + 0.02 * rng.standard_normal(Q), -0.99, 0.99)
                angles[2, q] = np.mod(chi0 + 0.05 * rng.standard_normal(Q), 2 * math.pi) if t.molecule == 2 else 0.0
            atom += 1
    return coords, angles


# ----------------------------------------------------------------------------
# the five BASELINE configurations
# ----------------------------------------------------------------------------
ROT_CONSTANTS = {            # A, B, C in cm^-1 used for the synthetic density-matrix tables
    "HCOOCH3": (0.6666525, 0.2306476, 0.1769383),   # nmv_prop/a-run:4
    "SO2": (2.02736, 0.34417, 0.29353),
    "H2O": (27.8806, 14.5216, 9.2778),
}


_deck_worm = {"C2": ("He4", 0.13, 16), "C3": ("H2", 0.35, 16)}      # WORM lines of the example decks


@dataclasses.dataclass
class Config:
    name: str
    system: System
    tables: Dict[str, object]
    coords: np.ndarray
    angles: np.ndarray
    perm: Optional[np.ndarray]
    deck_dir: str


def _deck(name: str) -> str:
    return os.path.join(DECKS, name)


def make_config(name: str, P: Optional[int] = None, Q: Optional[int] = None, nsolv: Optional[int] = None,
                seed: int = 1, big_tables: bool = True, temperature: Optional[float] = None, worm: bool = False) -> Config:
    """Build C1..C5.  P/Q/nsolv override the deck for reduced-size parity cases; worm=True keeps
    the deck's WORM line (C2: He4 0.13 16, C3: H2 0.35 16) so exchange is sampled with the worm algorithm."""
    tables: Dict[str, object] = {}
    perm = None
    keep_worm = worm
    if name == "C5":
        d = _deck("N2O_5pH2_0.5K_512_128")
        s = parse_qmc_input(os.path.join(d, "qmc.input"))
        s.worm = None
        if P: s.temperature *= 1024.0 / P       # reduced-size cases keep tau of the full configuration
        s.P, s.Q = P or 1024, Q or 128
        s.types[0].numb = nsolv if nsolv is not None else 100
        g1 = load_columns(os.path.join(d, "parah2.pot"), 2)
        tables["pot1d"] = (g1[0].copy(), g1[1].copy())
        tables["pot2d"] = synth_pot2d()
        tables["pot2d_delta"] = (0.005, 0.002)
        rot = load_columns(os.path.join(d, "N2O_T0.5t128.rot"), 4)
        tables["rotlin"] = tuple(np.ascontiguousarray(rot[i]) for i in range(4))
        coords, angles = cluster_config(s, seed)
    elif name in ("C1", "C2"):
        d = _deck("MF_8He_0.37K_512_128" if name == "C2" else "MF_1He_0.37K_512_128")
        s = parse_qmc_input(os.path.join(d, "qmc.input"))
        s.worm = None
        if P: s.temperature *= float(s.P) / P; s.P = P
        if Q: s.Q = Q
        if nsolv is not None: s.types[0].numb = nsolv
        g1 = load_columns(os.path.join(d, "helium.pot"), 2)
        tables["pot1d"] = (g1[0].copy(), g1[1].copy())
        if big_tables:
            tables["pot3d"] = (501, 181, 181, 4.0, 20.0, synth_pot3d(501, 181, 181))
            tables["rot3d"] = synth_rot3d(s.temperature, s.Q, *ROT_CONSTANTS["HCOOCH3"])
        full = (s.P == 512 and s.types[0].numb in (1, 8))
        if full:
            c8, a8, p8 = load_xyz_init(os.path.join(_deck("MF_8He_0.37K_512_128"), "xyz.init"), 9 * 512, 8)
            if name == "C2":
                coords, angles, perm = c8, a8, p8
            else:   # first He path + rotor path of the 8-He file
                idx = np.r_[0:512, 8 * 512:9 * 512]
                coords, angles = np.ascontiguousarray(c8[:, idx]), np.ascontiguousarray(a8[:, idx])
        else:
            coords, angles = cluster_config(s, seed)
    elif name == "C3":
        d = _deck("SO2_4pH2_0.37K_1024_256")
        s = parse_qmc_input(os.path.join(d, "qmc.input"))
        s.worm = None
        if P: s.temperature *= float(s.P) / P; s.P = P
        if Q: s.Q = Q
        if nsolv is not None: s.types[0].numb = nsolv
        g1 = load_columns(os.path.join(d, "isoH2H208.pot"), 2)
        tables["pot1d"] = (g1[0].copy(), g1[1].copy())
        if big_tables:
            tables["pot3d"] = (501, 181, 91, 4.0, 20.0, synth_pot3d(501, 181, 91))
            tables["rot3d"] = synth_rot3d(s.temperature, s.Q, *ROT_CONSTANTS["SO2"])
        if s.P == 1024 and s.types[0].numb == 4:
            coords, angles, perm = load_xyz_init(os.path.join(d, "xyz.init"), 5 * 1024, 4)
        else:
            coords, angles = cluster_config(s, seed)
    elif name == "C4":
        d = _deck("H2Odimer_0.74K_4096_2048")
        s = parse_qmc_input(os.path.join(d, "qmc.input"))
        s.rotden_type = 0
        if P: s.temperature *= float(s.P) / P; s.P = P
        if Q: s.Q = Q
        if big_tables:
            tables["rot3d"] = synth_rot3d(s.temperature, s.Q, *ROT_CONSTANTS["H2O"])
        coords, angles = cluster_config(s, seed)
    elif name == "SPH":
        # spherical treatment of a non-linear dopant (negative species count, mc_input.cc:152-156): the SO2 deck with
        # ROTATION removed (ISPHER is not compatible with it, mc_input.cc:463-464) and the vspher_ radial table
        d = _deck("SO2_4pH2_0.37K_1024_256")
        s = parse_qmc_input(os.path.join(d, "qmc.input"))
        s.worm = None
        s.ispher = 1
        s.Q = 0
        s.reflect, s.rotsym = (0, 0, 0), 0
        for t in s.types:
            t.rtstep = 0.0
        if P: s.temperature *= float(s.P) / P; s.P = P
        if nsolv is not None: s.types[0].numb = nsolv
        g1 = load_columns(os.path.join(d, "isoH2H208.pot"), 2)
        tables["pot1d"] = (g1[0].copy(), g1[1].copy())
        tables["vspher"] = load_vspher_table()
        coords, angles = cluster_config(s, seed)
    elif name == "CO2":
        # examples/CO2_100K_4_4: one free linear rotor (zero potential), the only deck of the reference whose files are all in the tree
        d = _deck("CO2_100K_4_4")
        s = parse_qmc_input(os.path.join(d, "qmc.input"))
        s.rotden_type = 0
        rot = load_columns(os.path.join(d, "CO2_T100t4.rot"), 4)
        tables["rotlin"] = tuple(np.ascontiguousarray(rot[i]) for i in range(4))
        coords, angles = cluster_config(s, seed)
    else:
        raise ValueError(name)
    if temperature is not None:
        s.temperature = temperature
    if keep_worm and _deck_worm.get(name):
        s.worm = _deck_worm[name]
    return Config(name, s, tables, coords, angles, perm, d)
