"""ORACLE bindings (test infrastructure, not product code).

ctypes wrappers over oracle/libpimcoracle.so (the travelling CPU restatement)
and, where it has been built, oracle/_ref/libpimcref.so (the reference's own
C++ objects).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libpimcoracle.so")
REF_SO = os.path.join(HERE, "_ref", "libpimcref.so")
REF_FAST_SO = os.path.join(HERE, "_ref", "libpimcref_fast.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


class OrcType(C.Structure):
    _fields_ = [("numb", C.c_int), ("molecule", C.c_int), ("stat", C.c_int), ("levels", C.c_int),
                ("mass", C.c_double), ("mcstep", C.c_double), ("rtstep", C.c_double)]


class OrcSystem(C.Structure):
    _fields_ = [("ntypes", C.c_int), ("type", OrcType * 2), ("P", C.c_int), ("Q", C.c_int),
                ("temperature", C.c_double), ("ispher", C.c_int), ("minimage", C.c_int), ("box", C.c_double * 3),
                ("rotden_type", C.c_int), ("rot_odevn", C.c_int), ("rot_eoff", C.c_double), ("x_rot", C.c_double),
                ("y_rot", C.c_double), ("z_rot", C.c_double), ("rnratio", C.c_int)]


def build_port(force: bool = False) -> str:
    if force or not os.path.exists(PORT_SO):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return PORT_SO


def system_struct(s) -> OrcSystem:
    o = OrcSystem()
    o.ntypes = len(s.types)
    for i, t in enumerate(s.types):
        o.type[i] = OrcType(t.numb, t.molecule, t.stat, t.levels, t.mass, t.mcstep, t.rtstep)
    o.P, o.Q, o.temperature = s.P, s.Q, s.temperature
    o.ispher, o.minimage = s.ispher, s.minimage
    n_atoms = sum(t.numb for t in s.types if t.molecule == 0)
    n_mols = sum(t.numb for t in s.types if t.molecule)
    box = ((n_atoms + n_mols) / s.density) ** (1.0 / 3.0)      # mc_setup.cc:339,359-360
    for d in range(3):
        o.box[d] = box
    o.rotden_type, o.rot_odevn, o.rot_eoff = s.rotden_type, s.rot_odevn, s.rot_eoff
    o.x_rot, o.y_rot, o.z_rot, o.rnratio = s.x_rot, s.y_rot, s.z_rot, s.rnratio
    return o


class Oracle:
    """CPU restatement handle; keeps numpy tables alive for the borrowed pointers."""

    def __init__(self, cfg):
        self.lib = L = C.CDLL(build_port())
        self.cfg = cfg
        s = cfg.system
        L.orc_create.restype = C.c_void_p
        self._sys = system_struct(s)
        self.h = C.c_void_p(L.orc_create(C.byref(self._sys)))
        for name in ("orc_spot1d", "orc_lpot2d", "orc_srotdens", "orc_vcord", "orc_caleng", "orc_pot_energy_it",
                     "orc_pot_energy_path", "orc_pot_rot_energy", "orc_pot_rot_e3d", "orc_get_kin", "orc_get_pot",
                     "orc_get_rot_energy", "orc_get_rot_e3d"):
            getattr(L, name).restype = C.c_double
        self._keep = []
        t = cfg.tables
        if "pot1d" in t:
            g, v = (np.ascontiguousarray(x, dtype=np.float64) for x in t["pot1d"])
            L.orc_set_pot1d(self.h, C.c_int(len(g)), _dp(g), _dp(v))
        if "pot2d" in t:
            rg, cg, v = (np.ascontiguousarray(x, dtype=np.float64) for x in t["pot2d"])
            dr, dc = float(rg[1] - rg[0]), float(cg[1] - cg[0])
            dr, dc = t.get("pot2d_delta", (round(dr, 12), round(dc, 12)))
            L.orc_set_pot2d(self.h, C.c_int(len(rg)), C.c_int(len(cg)), C.c_double(dr), C.c_double(dc), _dp(rg), _dp(cg), _dp(v))
        if "pot3d" in t:
            rg, thg, chg, rmin, rmax, v = t["pot3d"]
            self._keep.append(v)
            L.orc_set_pot3d(self.h, C.c_int(rg), C.c_int(thg), C.c_int(chg), C.c_double(rmin), C.c_double(rmax), _dp(v))
        if "rotlin" in t:
            a = [np.ascontiguousarray(x, dtype=np.float64) for x in t["rotlin"]]
            L.orc_set_rotlin(self.h, C.c_int(len(a[0])), _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]))
        if "rot3d" in t:
            self._keep.extend(t["rot3d"])
            L.orc_set_rot3d(self.h, *[_dp(x) for x in t["rot3d"]])
        self.N, self.P, self.Q = s.N, s.P, s.Q
        self.set_state(cfg.coords, cfg.angles, cfg.perm)
        if getattr(s, "worm", None):
            self.worm_init([t.name for t in s.types].index(s.worm[0]), s.worm[1], s.worm[2])

    def __del__(self):
        try:
            self.lib.orc_destroy(self.h)
        except Exception:
            pass

    # state ---------------------------------------------------------------
    def set_state(self, coords, angles, perm=None):
        c = np.ascontiguousarray(coords, dtype=np.float64)
        a = np.ascontiguousarray(angles, dtype=np.float64)
        p = np.ascontiguousarray(perm, dtype=np.int32) if perm is not None else None
        self.lib.orc_set_state(self.h, _dp(c), _dp(a), _ip(p))

    def get_state(self):
        n = self.N * self.P
        c, a, cs = (np.zeros((3, n)) for _ in range(3))
        self.lib.orc_get_state(self.h, _dp(c), _dp(a), _dp(cs))
        return c, a, cs

    # leaves ----------------------------------------------------------------
    def spot1d(self, r):
        k = C.c_int()
        v = self.lib.orc_spot1d(self.h, C.c_double(r), C.byref(k))
        return v, k.value

    def lpot2d(self, r, c):
        ir, ic = C.c_int(), C.c_int()
        v = self.lib.orc_lpot2d(self.h, C.c_double(r), C.c_double(c), C.byref(ir), C.byref(ic))
        return v, ir.value, ic.value

    def srotdens(self, g, which=0):
        return self.lib.orc_srotdens(self.h, C.c_double(g), C.c_int(which))

    def rotden(self, e1, e2):
        e1 = np.ascontiguousarray(e1, dtype=np.float64); e2 = np.ascontiguousarray(e2, dtype=np.float64)
        rel = np.zeros(3); rho, erot, esq = C.c_double(), C.c_double(), C.c_double()
        idx, istop = C.c_int(), C.c_int()
        self.lib.orc_rotden(self.h, _dp(e1), _dp(e2), _dp(rel), C.byref(rho), C.byref(erot), C.byref(esq), C.byref(idx), C.byref(istop))
        return rho.value, erot.value, esq.value, idx.value, rel

    def vcord(self, eul, rcom, rpt):
        eul, rcom, rpt = (np.ascontiguousarray(x, dtype=np.float64) for x in (eul, rcom, rpt))
        rtc = np.zeros(3); idx = C.c_int()
        v = self.lib.orc_vcord(self.h, _dp(eul), _dp(rcom), _dp(rpt), _dp(rtc), C.byref(idx))
        return v, rtc, idx.value

    def rotpro(self, deg3):
        """rotpro_sub.f on (phi, theta, chi) in degrees -> rho, erot, esq (table units), flat index, jstop"""
        d = np.ascontiguousarray(deg3, dtype=np.float64)
        rho, erot, esq = C.c_double(), C.c_double(), C.c_double(); idx, js = C.c_int(), C.c_int()
        self.lib.orc_rotpro(self.h, _dp(d), C.byref(rho), C.byref(erot), C.byref(esq), C.byref(idx), C.byref(js))
        return rho.value, erot.value, esq.value, idx.value, js.value

    def vcalc(self, rtc):
        """vcalc.f on (r bohr, theta deg, chi deg) -> V, flat index"""
        d = np.ascontiguousarray(rtc, dtype=np.float64); idx = C.c_int()
        self.lib.orc_vcalc.restype = C.c_double
        v = self.lib.orc_vcalc(self.h, _dp(d), C.byref(idx))
        return v, idx.value

    def deleul(self, e1, e2):
        e1 = np.ascontiguousarray(e1, dtype=np.float64); e2 = np.ascontiguousarray(e2, dtype=np.float64)
        rel = np.zeros(3)
        self.lib.orc_deleul(_dp(e1), _dp(e2), _dp(rel))
        return rel

    def vspher(self, r):
        """vspher_ -> (V, the clamped r in bohr the Fortran leaves in its argument)"""
        rc = C.c_double()
        self.lib.orc_vspher.restype = C.c_double
        v = self.lib.orc_vspher(C.c_double(r), C.byref(rc))
        return v, rc.value

    def caleng(self, c1, c2, e1, e2):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (c1, c2, e1, e2)]
        return self.lib.orc_caleng(*[_dp(x) for x in a])

    # sums --------------------------------------------------------------------
    def pot_energy_it(self, atom, it, pos=None):
        p = np.ascontiguousarray(pos, dtype=np.float64) if pos is not None else None
        return self.lib.orc_pot_energy_it(self.h, C.c_int(atom), _dp(p), C.c_int(it))

    def pot_energy_path(self, atom, shift=None):
        p = np.ascontiguousarray(shift, dtype=np.float64) if shift is not None else None
        return self.lib.orc_pot_energy_path(self.h, C.c_int(atom), _dp(p))

    def pot_rot_energy(self, atom, cos3, it):
        p = np.ascontiguousarray(cos3, dtype=np.float64)
        return self.lib.orc_pot_rot_energy(self.h, C.c_int(atom), _dp(p), C.c_int(it))

    def pot_rot_e3d(self, atom, eul, it):
        p = np.ascontiguousarray(eul, dtype=np.float64)
        return self.lib.orc_pot_rot_e3d(self.h, C.c_int(atom), _dp(p), C.c_int(it))

    # moves ---------------------------------------------------------------------
    def bisection_move(self, typ, atom, time, ug, ua, exch=0):
        ug = np.ascontiguousarray(ug, dtype=np.float64); ua = np.ascontiguousarray(ua, dtype=np.float64)
        used = (C.c_int * 2)()
        acc = self.lib.orc_bisection_move(self.h, typ, atom, time, _dp(ug), _dp(ua), exch, used)
        return acc, used[0], used[1]

    def molecular_move(self, typ, atom, u3, ua):
        u3 = np.ascontiguousarray(u3, dtype=np.float64)
        return self.lib.orc_molecular_move(self.h, typ, atom, _dp(u3), C.c_double(ua))

    def rot3d_step(self, it1, atom0, typ, r):
        return self.lib.orc_rot3d_step(self.h, it1, atom0, typ, *[C.c_double(x) for x in r])

    def rotlin_step(self, it1, typ, r):
        return self.lib.orc_rotlin_step(self.h, it1, typ, *[C.c_double(x) for x in r])

    # estimators ------------------------------------------------------------------
    def get_kin(self):
        return self.lib.orc_get_kin(self.h)

    def get_pot(self, dens=0):
        return self.lib.orc_get_pot(self.h, C.c_int(dens))

    def get_rot_energy(self):
        a, b = C.c_double(), C.c_double()
        f = self.lib.orc_get_rot_e3d if self.cfg.system.types[-1].molecule == 2 else self.lib.orc_get_rot_energy
        s = f(self.h, C.byref(a), C.byref(b))
        return s, a.value, b.value

    def get_rcf(self):
        out = np.zeros(self.Q)
        self.lib.orc_get_rcf(self.h, _dp(out))
        return out

    def reset_hist(self):
        self.lib.orc_reset_hist(self.h)

    def get_hist(self):
        g1, g2 = np.zeros(300), np.zeros(300 * 50)
        g3a, g3m = np.zeros(300 * 50 * 100), np.zeros(300 * 50 * 100)
        rt, rp, rc = np.zeros(50), np.zeros(100), np.zeros(100)
        self.lib.orc_get_hist(self.h, _dp(g1), _dp(g2), _dp(g3a), _dp(g3m), _dp(rt), _dp(rp), _dp(rc))
        return dict(gr1d=g1, gr2d=g2, gr3d_atoms=g3a, gr3d_mols=g3m, relthe=rt, relphi=rp, relchi=rc)

    # area / exchange estimators and symmetry operations ---------------------------------
    def exchange_length(self):
        nb = max(t.numb for t in self.cfg.system.types if t.stat == 1)
        out = np.zeros(nb)
        self.lib.orc_exchange_length(self.h, _dp(out))
        return out

    def area_estimators(self):
        out = np.zeros(4)
        self.lib.orc_area_estimators(self.h, _dp(out))
        return out

    def area_estim3d(self, iframe):
        a, i = np.zeros(3), np.zeros(9)
        self.lib.orc_area_estim3d(self.h, C.c_int(iframe), _dp(a), _dp(i))
        return a, i

    def reflect(self, plane):
        self.lib.orc_reflect(self.h, C.c_int(plane))

    def rotsym(self, u, nfold):
        self.lib.orc_rotsym(self.h, C.c_double(u), C.c_int(nfold))

    def sched_symmetry(self, refl_x, refl_y, refl_z, rotsym, nfold):
        self.lib.orc_sched_symmetry(self.h, *[C.c_int(v) for v in (refl_x, refl_y, refl_z, rotsym, nfold)])

    # worm ------------------------------------------------------------------------------
    def worm_init(self, type_, c_input, m):
        self.lib.orc_worm_init(self.h, C.c_int(type_), C.c_double(c_input), C.c_int(m))

    def worm_set(self, st5):
        self.lib.orc_worm_set(self.h, (C.c_int * 5)(*[int(v) for v in st5]))

    def worm_get(self):
        st = (C.c_int * 5)()
        self.lib.orc_worm_get(self.h, st)
        return list(st)

    def worm_push(self, stream, u):
        u = np.ascontiguousarray(np.atleast_1d(u), dtype=np.float64)
        self.lib.orc_worm_push(self.h, C.c_int(stream), _dp(u), C.c_int(len(u)))

    def worm_clear(self):
        self.lib.orc_worm_clear(self.h)

    def worm_pending(self, stream):
        return self.lib.orc_worm_pending(self.h, C.c_int(stream))

    def worm_op(self, which, sched_stream=False):
        self.lib.orc_worm_op(self.h, C.c_int(which), C.c_int(1 if sched_stream else 0))

    def worm_counters(self):
        t, a, cq = np.zeros(7), np.zeros(7), C.c_double()
        self.lib.orc_worm_counters(self.h, _dp(t), _dp(a), C.byref(cq))
        return t, a, cq.value

    def get_perm(self, n):
        p, r = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        self.lib.orc_get_perm(self.h, _ip(p), _ip(r), C.c_int(n))
        return p, r

    def world_line(self, atom, pt):
        return bool(self.lib.orc_world_line(self.h, C.c_int(atom), C.c_int(pt)))

    # schedule replay -----------------------------------------------------------------
    def sched_seed(self, seed6, chain_global=0):
        sd = (C.c_ulong * 6)(*seed6)
        self.lib.orc_sched_seed(self.h, sd, C.c_long(chain_global))

    def sched_run(self, t0, nsteps):
        self.lib.orc_sched_run(self.h, C.c_long(t0), C.c_long(nsteps))

    def counters(self):
        t, a = np.zeros(6), np.zeros(6)
        self.lib.orc_sched_counters(self.h, _dp(t), _dp(a))
        return t.reshape(2, 3), a.reshape(2, 3)


def mrg_draws(seed6, first_stream, nstream, ndraw):
    L = C.CDLL(build_port())
    out = np.zeros((nstream, ndraw))
    L.orc_mrg_draws((C.c_ulong * 6)(*seed6), C.c_long(first_stream), C.c_int(nstream), C.c_int(ndraw), _dp(out))
    return out


# ----------------------------------------------------------------------------
# the reference's own objects (only where oracle/_ref has been built)
# ----------------------------------------------------------------------------
def ref_available(fast: bool = False) -> bool:
    return os.path.exists(REF_FAST_SO if fast else REF_SO)


class Ref:
    """One reference instance per PROCESS (the reference keeps its state in globals)."""
    _used = False

    def __init__(self, cfg, nthreads=1, fast=False, workdir=None):
        if Ref._used:
            raise RuntimeError("the reference library holds global state: one Ref per process")
        Ref._used = True
        from importlib import import_module  # noqa: F401
        self.cfg = cfg
        s = cfg.system
        self.lib = L = C.CDLL(REF_FAST_SO if fast else REF_SO)
        self.work = workdir or tempfile.mkdtemp(prefix="pimcref_")
        os.makedirs(os.path.join(self.work, "out"), exist_ok=True)
        import sys
        sys.path.insert(0, os.path.dirname(HERE))
        cfgmod = _configs()
        cfgmod.write_qmc_input(s, os.path.join(self.work, "qmc.input"))
        t = cfg.tables
        for ty in s.types:
            if ty.molecule == 0:
                g, v = t["pot1d"]
                np.savetxt(os.path.join(self.work, ty.fpot + ".pot"), np.c_[g, v], fmt="%.17g")
            elif ty.molecule == 1:
                if "pot2d" in t:
                    rg, cg, v = t["pot2d"]
                    dr, dc = t.get("pot2d_delta", (round(float(rg[1] - rg[0]), 12), round(float(cg[1] - cg[0]), 12)))
                    cfgmod.write_pot2d(os.path.join(self.work, ty.fpot + ".pot"), rg, cg, v, dr, dc)
                else:       # a lone rotor: the reference's CO2_fake.pot
                    open(os.path.join(self.work, ty.fpot + ".pot"), "w").write("0 0\n0 0\n")
                if s.Q and s.rotden_type == 0:
                    # init_rotdens file name: type + "_T" + temperature + "t" + Q (mc_poten.cc:518-524)
                    fn = f"{ty.name}_T{_cxx_double(s.temperature)}t{s.Q}.rot"
                    np.savetxt(os.path.join(self.work, fn), np.c_[t["rotlin"][0], t["rotlin"][1], t["rotlin"][2], t["rotlin"][3]], fmt="%.17g")
        vt = t.get("pot3d")
        r3 = t.get("rot3d")
        self._keep = [vt, r3]
        L.ref_init.argtypes = [C.c_char_p, c_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, c_dp, c_dp, c_dp, C.c_int]
        for name in ("ref_SPot1D", "ref_LPot2D", "ref_SRotDens", "ref_SRotDensDeriv", "ref_SRotDensEsqrt", "ref_PotEnergy_it",
                     "ref_PotEnergy_path", "ref_PotRotEnergy", "ref_PotRotE3D", "ref_GetKinEnergy", "ref_GetPotEnergy",
                     "ref_GetPotEnergy_Densities", "ref_GetRotEnergy", "ref_GetRotE3D", "ref_run_steps", "ref_tau", "ref_lambda"):
            getattr(L, name).restype = C.c_double
        L.ref_SPot1D.argtypes = [C.c_double, C.c_int]
        L.ref_LPot2D.argtypes = [C.c_double, C.c_double, C.c_int]
        for n_ in ("ref_SRotDens", "ref_SRotDensDeriv", "ref_SRotDensEsqrt"):
            getattr(L, n_).argtypes = [C.c_double, C.c_int]
        L.ref_MCRot3Dstep.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_double] * 4
        L.ref_MCRotLinStep.argtypes = [C.c_int, C.c_int] + [C.c_double] * 3
        cwd = os.getcwd()
        rc = L.ref_init(self.work.encode(), _dp(vt[5]) if vt else None, *(vt[:3] if vt else (0, 0, 0)),
                        *(map(float, vt[3:5]) if vt else (0.0, 0.0)),
                        _dp(r3[0]) if r3 else None, _dp(r3[1]) if r3 else None, _dp(r3[2]) if r3 else None, nthreads)
        os.chdir(cwd)
        if rc:
            raise RuntimeError("ref_init failed")
        self.N, self.P, self.Q = s.N, s.P, s.Q
        self.set_state(cfg.coords, cfg.angles, cfg.perm)

    def set_state(self, coords, angles, perm=None):
        c = np.ascontiguousarray(coords, dtype=np.float64); a = np.ascontiguousarray(angles, dtype=np.float64)
        p = np.ascontiguousarray(perm, dtype=np.int32) if perm is not None else None
        self.lib.ref_set_state(_dp(c), _dp(a), _ip(p))

    def get_state(self):
        n = self.N * self.P
        c, a, cs = (np.zeros((3, n)) for _ in range(3))
        self.lib.ref_get_state(_dp(c), _dp(a), _dp(cs))
        return c, a, cs

    def push(self, stream, u):
        u = np.ascontiguousarray(np.atleast_1d(u), dtype=np.float64)
        self.lib.ref_rng_push(C.c_int(stream), _dp(u), C.c_int(len(u)))

    def queue_mode(self, on=True):
        self.lib.ref_rng_queue_mode(C.c_int(1 if on else 0))
        self.lib.ref_rng_clear()

    def rot_energy(self):
        a, b = C.c_double(), C.c_double()
        f = self.lib.ref_GetRotE3D if self.cfg.system.types[-1].molecule == 2 else self.lib.ref_GetRotEnergy
        s = f(C.byref(a), C.byref(b))
        return s, a.value, b.value


def _cxx_double(x: float) -> str:
    """ostream << double with the default precision 6 (how init_rot3D/init_rotdens build file names)."""
    return "%g" % x


def _configs():
    import importlib.util
    import sys
    name = "moribs_pimc_b200"
    if name in sys.modules:
        return sys.modules[name].configs
    root = os.path.dirname(HERE)
    spec = importlib.util.spec_from_file_location(name, os.path.join(root, "moribs-pimc_b200", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(root, "moribs-pimc_b200")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod.configs
