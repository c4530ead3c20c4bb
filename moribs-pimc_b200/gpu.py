"""ctypes binding of libpimcgpu.so (include/pimcgpu.h) -- the Python host-side mirror used by the
tests and bench.py.  There is NO CPU fallback: if the CUDA library is missing or no GPU is
visible every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("PIMCGPU_LIB", os.path.join(CSRC, "libpimcgpu.so"))   # override only for kernel-variant experiments
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fopenmp", "-shared"]

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)

EXPORTS = [
    "pimcgpu_init", "pimcgpu_finalize", "pimcgpu_last_error", "pimcgpu_upload_state", "pimcgpu_download_state",
    "pimcgpu_seed", "pimcgpu_steps", "pimcgpu_sync", "pimcgpu_step_counter", "pimcgpu_geometry", "pimcgpu_measure", "pimcgpu_accum_layout",
    "pimcgpu_accum_device_ptr", "pimcgpu_accum_download", "pimcgpu_accum_download_begin", "pimcgpu_accum_download_end", "pimcgpu_accum_reset", "pimcgpu_block_scalars",
    "pimcgpu_counters", "pimcgpu_stream", "pimcgpu_chain_energies", "pimcgpu_chain_rcf", "pimcgpu_eval_spot1d",
    "pimcgpu_eval_lpot2d", "pimcgpu_eval_srotdens", "pimcgpu_eval_rotden", "pimcgpu_eval_vcord", "pimcgpu_eval_caleng",
    "pimcgpu_pot_energy_slice", "pimcgpu_rng_draws", "pimcgpu_fp64_peak", "pimcgpu_host_spline",
    "pimcgpu_host_stream_state", "pimcgpu_host_lut", "pimcgpu_accum_offset", "pimcgpu_symmetry_moves", "pimcgpu_symmetry_ops",
    "pimcgpu_checkpoint_bytes", "pimcgpu_checkpoint_save", "pimcgpu_checkpoint_load", "pimcgpu_chain_areas", "pimcgpu_worm_moves", "pimcgpu_worm_state", "pimcgpu_worm_set", "pimcgpu_worm_counters",
    "pimcgpu_upload_states", "pimcgpu_download_states", "pimcgpu_download_states_rows",
    "pimcgpu_upload_states_begin", "pimcgpu_upload_states_commit", "pimcgpu_download_states_begin", "pimcgpu_download_states_end",
    "pimcgpu_gen_asymrho", "pimcgpu_gen_symrho", "pimcgpu_gen_linden", "pimcgpu_gen_wigner_d", "pimcgpu_gen_timing",
    "pimcgpu_format_e15_8", "pimcgpu_write_e15_8", "pimcgpu_write_rot",
    "pimcgpu_eval_rotpro", "pimcgpu_eval_vcalc", "pimcgpu_eval_deleul", "pimcgpu_eval_vcord_grid", "pimcgpu_eval_vspher", "pimcgpu_eval_libm",
]


class GpuType(C.Structure):
    _fields_ = [("numb", C.c_int), ("molecule", C.c_int), ("stat", C.c_int), ("levels", C.c_int),
                ("mass", C.c_double), ("mcstep", C.c_double), ("rtstep", C.c_double)]


class GpuSystem(C.Structure):
    _fields_ = [("ntypes", C.c_int), ("type", GpuType * 2), ("P", C.c_int), ("Q", C.c_int), ("temperature", C.c_double),
                ("ispher", C.c_int), ("minimage", C.c_int), ("box", C.c_double * 3), ("rotden_type", C.c_int),
                ("nchains", C.c_int), ("chain_offset", C.c_long), ("device", C.c_int), ("ctas_per_chain", C.c_int),
                ("threads_per_cta", C.c_int), ("team", C.c_int),
                ("rot_odevn", C.c_int), ("rot_eoff", C.c_double), ("x_rot", C.c_double), ("y_rot", C.c_double), ("z_rot", C.c_double),
                ("rnratio", C.c_int), ("reflect", C.c_int * 3), ("rotsym", C.c_int), ("nfold_rot", C.c_int),
                ("worm", C.c_int), ("worm_type", C.c_int), ("worm_c", C.c_double), ("worm_m", C.c_int)]


class GpuTables(C.Structure):
    _fields_ = [("n1d", C.c_int), ("grid1d", c_dp), ("pot1d", c_dp),
                ("rsize2d", C.c_int), ("csize2d", C.c_int), ("dr2d", C.c_double), ("dc2d", C.c_double),
                ("rgrid2d", c_dp), ("cgrid2d", c_dp), ("pot2d", c_dp),
                ("rgrd", C.c_int), ("thgrd", C.c_int), ("chgrd", C.c_int), ("rvmin", C.c_double), ("rvmax", C.c_double),
                ("vtable", c_dp),
                ("nrot", C.c_int), ("rotgrid", c_dp), ("rotdens", c_dp), ("rotderv", c_dp), ("rotesqr", c_dp),
                ("rho3d", c_dp), ("erot3d", c_dp), ("esq3d", c_dp), ("vspher", c_dp)]


class GpuScalars(C.Structure):
    _fields_ = [("count", C.c_double), ("kin", C.c_double), ("pot", C.c_double), ("rot", C.c_double), ("rotsq", C.c_double),
                ("cv", C.c_double), ("cv_trans", C.c_double), ("cv_rot", C.c_double),
                ("mctotal", (C.c_double * 3) * 2), ("mcaccep", (C.c_double * 3) * 2)]


def build(force: bool = False) -> str:
    """nvcc cross-compile of the in-tree library for sm_100a (works without a GPU): one object per translation unit
    (pimcgpu.cu = the sampling path, pimc_tablegen.cu = the rho-table generators), linked into libpimcgpu.so."""
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "pimcgpu.h"))
    units = {"pimcgpu.cu": hdrs, "pimc_tablegen.cu": hdrs[-1:]}
    objs, relink = [], force or not os.path.exists(LIB)
    for cu, deps in units.items():
        src, obj = os.path.join(CSRC, cu), os.path.join(CSRC, cu[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or any(os.path.getmtime(obj) < os.path.getmtime(d) for d in [src] + deps):
            subprocess.check_call(["nvcc"] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", "-o", obj, src])
            relink = True
    if relink or any(os.path.getmtime(LIB) < os.path.getmtime(o) for o in objs):
        subprocess.check_call(["nvcc", "-shared", "-Xcompiler", "-fopenmp", "-o", LIB] + objs)
    return LIB


_lib = None


def lib():
    """Load libpimcgpu.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(f"{LIB} is missing: run __graft_entry__.build() (nvcc, sm_100a) first; there is no CPU fallback")
        L = C.CDLL(LIB)
        L.pimcgpu_last_error.restype = C.c_char_p
        L.pimcgpu_step_counter.restype = C.c_long
        L.pimcgpu_accum_device_ptr.restype = C.c_void_p
        L.pimcgpu_stream.restype = C.c_void_p
        L.pimcgpu_steps.argtypes = [C.c_long]
        L.pimcgpu_rng_draws.argtypes = [C.c_long, C.c_int, c_dp]
        L.pimcgpu_accum_download.argtypes = [c_dp, C.c_long]
        L.pimcgpu_accum_download_begin.argtypes = [c_dp, C.c_long]
        L.pimcgpu_accum_offset.restype = C.c_long
        L.pimcgpu_checkpoint_bytes.restype = C.c_long
        L.pimcgpu_accum_offset.argtypes = [C.c_char_p]
        L.pimcgpu_format_e15_8.restype = None
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


def eval_libm(which, x):
    """device libm: which = sin, cos, acos, atan, exp, log, sqrt, fmod2pi"""
    names = ("sin", "cos", "acos", "atan", "exp", "log", "sqrt", "fmod2pi")
    x = np.ascontiguousarray(x, dtype=np.float64); y = np.zeros_like(x)
    _ck(lib().pimcgpu_eval_libm(C.c_int(names.index(which)), C.c_int(len(x)), _dp(x), _dp(y)))
    return y


def fp64_peak_tflops() -> float:
    v = C.c_double()
    _ck(lib().pimcgpu_fp64_peak(C.byref(v)))
    return v.value


class PimcGpuError(RuntimeError):
    pass


def _ck(rc):
    if rc:
        raise PimcGpuError(lib().pimcgpu_last_error().decode())


# ---- rho-table generators (csrc/pimc_tablegen.cu; no context needed) -------------------------------------
NPLANE = 361 * 361


def _planes(nt):
    return [np.zeros((nt, 361, 361)) for _ in range(3)]


def gen_asymrho(T, nslice, iodevn, ith0, ith1, A, B, Cc, maxj):
    """asymrho.x T P iodevn ith0 ithend A B C maxj -> rho, eng, esq [ntheta][361][361], info[16]"""
    r, e, q = _planes(ith1 - ith0 + 1)
    info = np.zeros(16)
    _ck(lib().pimcgpu_gen_asymrho(C.c_double(T), C.c_int(nslice), C.c_int(iodevn), C.c_int(ith0), C.c_int(ith1), C.c_double(A),
                                  C.c_double(B), C.c_double(Cc), C.c_int(maxj), _dp(r), _dp(e), _dp(q), _dp(info)))
    return r, e, q, info


def asym_auto_maxj(T, nslice, A, B, Cc):
    """first j whose lowest level falls under the generator's own cut (2j+1)/8pi^2 e^{-tau E} < 1e-16 (asymrho.f:520);
    the rule pimc_b200 uses when it generates missing tables"""
    tau = 1.0 / (0.6950356 * T) / nslice
    cmin = min(A, B, Cc)
    j = 4
    while j < 876 and (2 * j + 1) / (8.0 * np.pi ** 2) * np.exp(-tau * cmin * j * (j + 1.0)) >= 1e-16:
        j += 1
    return j


def gen_timing():
    ms = np.zeros(4)
    _ck(lib().pimcgpu_gen_timing(_dp(ms)))
    return ms


def gen_symrho(T, nslice, kmod, ith0, ith1, Bz, Bxy, maxj):
    """symrho.x T P kmod ith0 ithend Bz Bxy maxj -> rho, eng, esq [ntheta][361][361], info[5]"""
    r, e, q = _planes(ith1 - ith0 + 1)
    info = np.zeros(5)
    _ck(lib().pimcgpu_gen_symrho(C.c_double(T), C.c_int(nslice), C.c_int(kmod), C.c_int(ith0), C.c_int(ith1), C.c_double(Bz),
                                 C.c_double(Bxy), C.c_int(maxj), _dp(r), _dp(e), _dp(q), _dp(info)))
    return r, e, q, info


def gen_linden(T, nslice, bconst, npt, iodevn):
    """linden.x T P B npt iodevn -> out[npt][4] (cos gamma, rho, erot, erotsq), info[4]"""
    out = np.zeros((npt, 4))
    info = np.zeros(4)
    _ck(lib().pimcgpu_gen_linden(C.c_double(T), C.c_int(nslice), C.c_double(bconst), C.c_int(npt), C.c_int(iodevn), _dp(out), _dp(info)))
    return out, info


def gen_wigner_d(maxj, theta):
    d = np.zeros((maxj + 1, 2 * maxj + 1, 2 * maxj + 1))
    _ck(lib().pimcgpu_gen_wigner_d(C.c_int(maxj), C.c_double(theta), _dp(d)))
    return d


def format_e15_8(v, scale1p=False):
    b = C.create_string_buffer(16)
    lib().pimcgpu_format_e15_8(C.c_double(v), C.c_int(1 if scale1p else 0), b)
    return b.value.decode()


def write_e15_8(path, v, append=False):
    a = np.ascontiguousarray(v, dtype=np.float64).ravel()
    _ck(lib().pimcgpu_write_e15_8(path.encode(), _dp(a), C.c_long(a.size), C.c_int(1 if append else 0)))


def write_rot(path, out4):
    a = np.ascontiguousarray(out4, dtype=np.float64)
    _ck(lib().pimcgpu_write_rot(path.encode(), _dp(a), C.c_int(a.shape[0])))


class PimcGpu:
    """One device context (the library is a per-process singleton, like the reference's globals)."""

    def __init__(self, cfg, nchains=1, chain_offset=0, device=0, ctas_per_chain=0, threads_per_cta=0, team=0):
        self.cfg = cfg
        s = cfg.system
        self.s = s
        self.N, self.P, self.Q = s.N, s.P, s.Q
        self.nchains = nchains
        sy = GpuSystem()
        sy.ntypes = len(s.types)
        for i, t in enumerate(s.types):
            sy.type[i] = GpuType(t.numb, t.molecule, t.stat, t.levels, t.mass, t.mcstep, t.rtstep)
        sy.P, sy.Q, sy.temperature = s.P, s.Q, s.temperature
        sy.ispher, sy.minimage, sy.rotden_type = s.ispher, s.minimage, s.rotden_type
        box = (s.N / s.density) ** (1.0 / 3.0)               # MCInit, mc_setup.cc:339,359-360
        for d in range(3):
            sy.box[d] = box
        sy.nchains, sy.chain_offset, sy.device = nchains, chain_offset, device
        sy.ctas_per_chain, sy.threads_per_cta, sy.team = ctas_per_chain, threads_per_cta, team
        sy.rot_odevn, sy.rot_eoff, sy.rnratio = s.rot_odevn, s.rot_eoff, s.rnratio
        sy.x_rot, sy.y_rot, sy.z_rot = s.x_rot, s.y_rot, s.z_rot
        for d in range(3):
            sy.reflect[d] = s.reflect[d]
        sy.rotsym, sy.nfold_rot = (1 if s.rotsym else 0), max(1, s.rotsym)      # ROTSYM n: IROTSYM = 1, NFOLD_ROT = n
        if s.worm:
            names = [t.name for t in s.types]
            sy.worm, sy.worm_type, sy.worm_c, sy.worm_m = 1, names.index(s.worm[0]), float(s.worm[1]), int(s.worm[2])
        tb = GpuTables()
        self._keep = []
        t = cfg.tables

        def arr(x):
            a = np.ascontiguousarray(x, dtype=np.float64)
            self._keep.append(a)
            return a

        if "pot1d" in t:
            g, v = arr(t["pot1d"][0]), arr(t["pot1d"][1])
            tb.n1d, tb.grid1d, tb.pot1d = len(g), _dp(g), _dp(v)
        if "pot2d" in t:
            rg, cg, v = (arr(x) for x in t["pot2d"])
            dr, dc = t.get("pot2d_delta", (round(float(rg[1] - rg[0]), 12), round(float(cg[1] - cg[0]), 12)))
            tb.rsize2d, tb.csize2d, tb.dr2d, tb.dc2d = len(rg), len(cg), dr, dc
            tb.rgrid2d, tb.cgrid2d, tb.pot2d = _dp(rg), _dp(cg), _dp(v)
        if "pot3d" in t:
            rg, thg, chg, rmin, rmax, v = t["pot3d"]
            v = arr(v)
            tb.rgrd, tb.thgrd, tb.chgrd, tb.rvmin, tb.rvmax, tb.vtable = rg, thg, chg, rmin, rmax, _dp(v)
        if "rotlin" in t:
            a = [arr(x) for x in t["rotlin"]]
            tb.nrot, tb.rotgrid, tb.rotdens, tb.rotderv, tb.rotesqr = len(a[0]), _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3])
        if "rot3d" in t:
            a = [arr(x) for x in t["rot3d"]]
            tb.rho3d, tb.erot3d, tb.esq3d = _dp(a[0]), _dp(a[1]), _dp(a[2])
        if "vspher" in t:
            tb.vspher = _dp(arr(t["vspher"]))
        self.L = lib()
        _ck(self.L.pimcgpu_init(C.byref(sy), C.byref(tb)))
        self.upload(-1, cfg.coords, cfg.angles, cfg.perm)

    def close(self):
        self.L.pimcgpu_finalize()

    # state -------------------------------------------------------------------------------
    def upload(self, chain, coords, angles, perm=None):
        c = np.ascontiguousarray(coords, dtype=np.float64)
        a = np.ascontiguousarray(angles, dtype=np.float64)
        p = np.ascontiguousarray(perm, dtype=np.int32) if perm is not None else None
        _ck(self.L.pimcgpu_upload_state(C.c_int(chain), _dp(c), _dp(a), _ip(p)))

    def download(self, chain=0):
        n = self.N * self.P
        c, a, cs = (np.zeros((3, n)) for _ in range(3))
        _ck(self.L.pimcgpu_download_state(C.c_int(chain), _dp(c), _dp(a), _dp(cs), None))
        return c, a, cs

    def download_into(self, chain, coords, angles, cosine=None):
        """download_state into caller-owned arrays [3][N*P] (e.g. pinned host memory that lives across steps, like the
        reference's MCCoords / MCAngles / MCCosine which are allocated once, mc_setup.cc:135-163)"""
        _ck(self.L.pimcgpu_download_state(C.c_int(chain), _dp(coords), _dp(angles), _dp(cosine), None))

    def upload_all(self, coords, angles, perm=None, first=0):
        """batched upload: coords/angles [count][3][N*P] (contiguous), perm [count][nb] or one permutation for all chains or None"""
        c = np.ascontiguousarray(coords, dtype=np.float64)
        a = np.ascontiguousarray(angles, dtype=np.float64)
        count = c.shape[0]
        p = None
        if perm is not None:
            p = np.ascontiguousarray(perm, dtype=np.int32)
            if p.ndim == 1:
                p = np.ascontiguousarray(np.tile(p, (count, 1)))
        _ck(self.L.pimcgpu_upload_states(C.c_int(first), C.c_int(count), _dp(c), _dp(a), _ip(p)))

    def download_all_into(self, coords, angles, cosine=None, first=0):
        """batched download into caller-owned contiguous arrays [count][3][N*P]"""
        _ck(self.L.pimcgpu_download_states(C.c_int(first), C.c_int(coords.shape[0]), _dp(coords), _dp(angles), _dp(cosine)))

    def download_rows_into(self, coords, angles, cosine=None, first=0):
        """batched download that writes only the rotor rows of angles / cosine (the caller's arrays live across steps)"""
        _ck(self.L.pimcgpu_download_states_rows(C.c_int(first), C.c_int(coords.shape[0]), _dp(coords), _dp(angles), _dp(cosine)))

    # split-phase transfers (the arrays must be contiguous float64 / int32 and stay alive and unchanged until commit / end)
    def upload_begin(self, coords, angles, perm=None, first=0):
        assert coords.flags["C_CONTIGUOUS"] and angles.flags["C_CONTIGUOUS"] and coords.dtype == np.float64
        self._up_keep = (coords, angles, perm)
        _ck(self.L.pimcgpu_upload_states_begin(C.c_int(first), C.c_int(coords.shape[0]), _dp(coords), _dp(angles), _ip(perm)))

    def upload_commit(self):
        _ck(self.L.pimcgpu_upload_states_commit())
        self._up_keep = None

    def download_begin(self, coords, angles, cosine=None, first=0):
        self._down_keep = (coords, angles, cosine)
        _ck(self.L.pimcgpu_download_states_begin(C.c_int(first), C.c_int(coords.shape[0]), _dp(coords), _dp(angles), _dp(cosine)))

    def download_end(self):
        _ck(self.L.pimcgpu_download_states_end())
        self._down_keep = None

    def seed(self, seed6=(12345,) * 6):
        _ck(self.L.pimcgpu_seed((C.c_ulong * 6)(*seed6)))

    # moves / estimators ----------------------------------------------------------------------
    def steps(self, n, sync=True):
        _ck(self.L.pimcgpu_steps(C.c_long(n)))
        if sync:
            _ck(self.L.pimcgpu_sync())

    def sync(self):
        _ck(self.L.pimcgpu_sync())

    def geometry(self):
        g = (C.c_int * 8)()
        _ck(self.L.pimcgpu_geometry(g))
        k = ("ctas_per_chain", "threads_per_cta", "team", "rot_group", "smem_bytes", "kind", "max_active_clusters", "nchains")
        return dict(zip(k, list(g)))

    def measure(self):
        _ck(self.L.pimcgpu_measure())

    def chain_energies(self, chain=0):
        out = np.zeros(5)
        _ck(self.L.pimcgpu_chain_energies(C.c_int(chain), _dp(out)))
        return dict(kin=out[0], pot=out[1], rot=out[2], erotsq=out[3], eterm=out[4])

    def chain_rcf(self, chain=0):
        out = np.zeros(self.Q)
        _ck(self.L.pimcgpu_chain_rcf(C.c_int(chain), _dp(out)))
        return out

    def accum_layout(self):
        v = [C.c_long() for _ in range(7)]
        _ck(self.L.pimcgpu_accum_layout(*[C.byref(x) for x in v]))
        k = ("n_total", "scalars", "gr1d", "gr2d", "gr3d", "rcf", "relbins")
        lay = dict(zip(k, [x.value for x in v]))
        for name in ("area", "ploops"):
            lay[name] = self.L.pimcgpu_accum_offset(name.encode())
        return lay

    def chain_areas(self, chain=0):
        """GetAreaEstimators / GetAreaEstim3D sums of one chain (a18)."""
        out = np.zeros(28)
        _ck(self.L.pimcgpu_chain_areas(C.c_int(chain), _dp(out)))
        return dict(lin=out[0:4], sff_area=out[4:7], sff_inert=out[7:16], mff_area=out[16:19], mff_inert=out[19:28])

    # checkpoint (N4) ---------------------------------------------------------------------------
    def checkpoint_save(self):
        n = self.L.pimcgpu_checkpoint_bytes()
        buf = (C.c_char * n)()
        _ck(self.L.pimcgpu_checkpoint_save(buf, C.c_long(n)))
        return bytes(buf)

    def checkpoint_load(self, blob):
        _ck(self.L.pimcgpu_checkpoint_load(C.c_char_p(blob), C.c_long(len(blob))))

    # worm --------------------------------------------------------------------------------------
    def worm_moves(self, sync=True):
        _ck(self.L.pimcgpu_worm_moves())
        if sync:
            _ck(self.L.pimcgpu_sync())

    def worm_state(self, chain=0):
        st = (C.c_int * 5)()
        _ck(self.L.pimcgpu_worm_state(C.c_int(chain), st))
        return list(st)

    def worm_set(self, chain, st5):
        _ck(self.L.pimcgpu_worm_set(C.c_int(chain), (C.c_int * 5)(*[int(v) for v in st5])))

    def worm_counters(self):
        t, a, cq = np.zeros(7), np.zeros(7), C.c_double()
        _ck(self.L.pimcgpu_worm_counters(_dp(t), _dp(a), C.byref(cq)))
        return t, a, cq.value

    def download_perm(self, chain=0):
        nb = max((t.numb for t in self.s.types if t.stat == 1), default=0)
        p = np.zeros(max(1, nb), dtype=np.int32)
        n = self.N * self.P
        c = np.zeros((3, n))
        _ck(self.L.pimcgpu_download_state(C.c_int(chain), _dp(c), None, None, _ip(p)))
        return p[:nb]

    def symmetry_moves(self):
        _ck(self.L.pimcgpu_symmetry_moves())

    def symmetry_ops(self, ops):
        """ops[chain][4] = XZ, YZ, XY reflection flags, rotor index of the symmetry rotation or -1."""
        o = np.ascontiguousarray(ops, dtype=np.int32).reshape(self.nchains, 4)
        _ck(self.L.pimcgpu_symmetry_ops(_ip(o)))

    def accum_download(self):
        lay = self.accum_layout()
        self.L.pimcgpu_accum_device_ptr()          # folds the move counters into the buffer
        out = np.zeros(lay["n_total"])
        _ck(self.L.pimcgpu_accum_download(_dp(out), C.c_long(len(out))))
        return out, lay

    def accum_download_into(self, out):
        """accumulator buffer into a caller-owned array (e.g. pinned, reused every block)"""
        self.L.pimcgpu_accum_device_ptr()          # folds the move counters into the buffer
        _ck(self.L.pimcgpu_accum_download(_dp(out), C.c_long(len(out))))

    def accum_download_begin(self, out):
        """queue the copy of the accumulator buffer into a caller-owned (pinned) array; accum_download_end() waits for it.
        The buffer is taken as it stands (pimcgpu_accum_device_ptr folded the move counters in, an all-reduce may have followed)."""
        _ck(self.L.pimcgpu_accum_download_begin(_dp(out), C.c_long(len(out))))

    def accum_download_end(self):
        _ck(self.L.pimcgpu_accum_download_end())

    def accum_reset(self):
        _ck(self.L.pimcgpu_accum_reset())

    def block_scalars(self):
        self.L.pimcgpu_accum_device_ptr()
        s = GpuScalars()
        _ck(self.L.pimcgpu_block_scalars(C.byref(s)))
        return s

    def counters(self):
        t, a = np.zeros(6), np.zeros(6)
        _ck(self.L.pimcgpu_counters(_dp(t), _dp(a)))
        return t.reshape(2, 3), a.reshape(2, 3)

    # parity entry points ------------------------------------------------------------------------
    def eval_spot1d(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64); v = np.zeros_like(r); k = np.zeros(len(r), dtype=np.int32)
        _ck(self.L.pimcgpu_eval_spot1d(len(r), _dp(r), _dp(v), _ip(k)))
        return v, k

    def eval_lpot2d(self, r, c):
        r = np.ascontiguousarray(r, dtype=np.float64); c = np.ascontiguousarray(c, dtype=np.float64)
        v = np.zeros_like(r); ir = np.zeros(len(r), dtype=np.int32); ic = np.zeros(len(r), dtype=np.int32)
        _ck(self.L.pimcgpu_eval_lpot2d(len(r), _dp(r), _dp(c), _dp(v), _ip(ir), _ip(ic)))
        return v, ir, ic

    def eval_srotdens(self, g, which=0):
        g = np.ascontiguousarray(g, dtype=np.float64); v = np.zeros_like(g)
        _ck(self.L.pimcgpu_eval_srotdens(len(g), _dp(g), C.c_int(which), _dp(v)))
        return v

    def eval_rotden(self, e1, e2):
        e1 = np.ascontiguousarray(e1, dtype=np.float64); e2 = np.ascontiguousarray(e2, dtype=np.float64)
        n = len(e1)
        rho, erot, esq = np.zeros(n), np.zeros(n), np.zeros(n); idx = np.zeros(n, dtype=np.int32)
        _ck(self.L.pimcgpu_eval_rotden(n, _dp(e1), _dp(e2), _dp(rho), _dp(erot), _dp(esq), _ip(idx)))
        return rho, erot, esq, idx

    def eval_vcord(self, eul, rcom, rpt):
        eul, rcom, rpt = (np.ascontiguousarray(x, dtype=np.float64) for x in (eul, rcom, rpt))
        n = len(eul)
        v = np.zeros(n); rtc = np.zeros((n, 3)); idx = np.zeros(n, dtype=np.int32)
        _ck(self.L.pimcgpu_eval_vcord(n, _dp(eul), _dp(rcom), _dp(rpt), _dp(v), _dp(rtc), _ip(idx)))
        return v, rtc, idx

    def eval_caleng(self, c1, c2, e1, e2):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (c1, c2, e1, e2)]
        n = len(a[0]); e = np.zeros(n)
        _ck(self.L.pimcgpu_eval_caleng(n, *[_dp(x) for x in a], _dp(e)))
        return e

    def eval_rotpro(self, deg):
        """rotpro on (phi, theta, chi) in degrees [n][3] -> rho, erot, esq (table units), flat index"""
        d = np.ascontiguousarray(deg, dtype=np.float64); n = len(d)
        rho, erot, esq = np.zeros(n), np.zeros(n), np.zeros(n); idx = np.zeros(n, dtype=np.int32)
        _ck(self.L.pimcgpu_eval_rotpro(n, _dp(d), _dp(rho), _dp(erot), _dp(esq), _ip(idx)))
        return rho, erot, esq, idx

    def eval_vcalc(self, rtc):
        """vcalc on (r bohr, theta deg, chi deg) [n][3] -> V, flat index"""
        d = np.ascontiguousarray(rtc, dtype=np.float64); n = len(d)
        v = np.zeros(n); idx = np.zeros(n, dtype=np.int32)
        _ck(self.L.pimcgpu_eval_vcalc(n, _dp(d), _dp(v), _ip(idx)))
        return v, idx

    def eval_deleul(self, e1, e2):
        e1 = np.ascontiguousarray(e1, dtype=np.float64); e2 = np.ascontiguousarray(e2, dtype=np.float64)
        rel = np.zeros((len(e1), 3))
        _ck(self.L.pimcgpu_eval_deleul(len(e1), _dp(e1), _dp(e2), _dp(rel)))
        return rel

    def eval_vcord_grid(self, eul, rcom, rpt):
        eul, rcom, rpt = (np.ascontiguousarray(x, dtype=np.float64) for x in (eul, rcom, rpt))
        g = np.zeros((len(eul), 3))
        _ck(self.L.pimcgpu_eval_vcord_grid(len(eul), _dp(eul), _dp(rcom), _dp(rpt), _dp(g)))
        return g

    def eval_vspher(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64); v = np.zeros_like(r); rc = np.zeros_like(r)
        _ck(self.L.pimcgpu_eval_vspher(len(r), _dp(r), _dp(v), _dp(rc)))
        return v, rc

    def pot_energy_slice(self, chain=0):
        v = np.zeros((self.N, self.P))
        _ck(self.L.pimcgpu_pot_energy_slice(C.c_int(chain), _dp(v)))
        return v

    def rng_draws(self, stream, n):
        out = np.zeros(n)
        _ck(self.L.pimcgpu_rng_draws(C.c_long(stream), C.c_int(n), _dp(out)))
        return out
