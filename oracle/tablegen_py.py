"""ORACLE bindings for the rho-table generators (test infrastructure, not product code).

ctypes wrappers over oracle/libtablegen_oracle.so, the CPU restatement of nmv_prop/asymrho.f,
symtop_prop/symrho.f and linear_prop/linden.f (real*16 parts in __float128).  Only tests/ and the
fixture scripts import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libtablegen_oracle.so")
c_dp = C.POINTER(C.c_double)
_lib = None


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            subprocess.check_call(["make", "-C", HERE, "libtablegen_oracle.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(SO)
        L.tg_wigd.restype = C.c_double
        L.tg_wigd.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double]
        L.tg_asym_setup.restype = C.c_void_p
        L.tg_asym_setup.argtypes = [C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, c_dp, C.c_char_p, C.c_int]
        L.tg_asym_free.argtypes = [C.c_void_p]
        L.tg_asym_nstates.argtypes = [C.c_void_p, C.c_int]
        L.tg_asym_energies.argtypes = [C.c_void_p, C.c_int, c_dp]
        L.tg_asym_dlist.argtypes = [C.c_void_p, C.c_int]
        L.tg_asym_point.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp]
        L.tg_asym_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp, c_dp]
        L.tg_asym_symfill.argtypes = [c_dp]
        L.tg_symrho_plane.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, c_dp, c_dp, c_dp, c_dp]
        L.tg_linden.argtypes = [C.c_double, C.c_int, C.c_double, C.c_int, C.c_int, c_dp, c_dp]
        L.tg_fmt_e15_8.argtypes = [C.c_double, C.c_int, C.c_char_p]
        _lib = L
    return _lib


def wigd(j, m, k, theta):
    return lib().tg_wigd(j, m, k, theta)


class AsymRho:
    """asymrho.f for one argument list; raises where the Fortran would STOP."""

    def __init__(self, T, nslice, iodevn, A, B, Cc, maxj):
        self.info = np.zeros(16)
        err = C.create_string_buffer(256)
        self.h = lib().tg_asym_setup(T, nslice, iodevn, A, B, Cc, maxj, _dp(self.info), err, 256)
        if not self.h:
            raise RuntimeError(err.value.decode())

    def close(self):
        if self.h:
            lib().tg_asym_free(self.h)
            self.h = None

    def energies(self, parity):
        e = np.zeros(lib().tg_asym_nstates(self.h, parity))
        lib().tg_asym_energies(self.h, parity, _dp(e))
        return e

    def point(self, ithe, iphi, ichi):
        o = np.zeros(3)
        lib().tg_asym_point(self.h, ithe, iphi, ichi, _dp(o))
        return o

    def plane(self, ithe, stride=1):
        r, e, q = (np.zeros((361, 361)) for _ in range(3))
        lib().tg_asym_plane(self.h, ithe, stride, _dp(r), _dp(e), _dp(q))
        return r, e, q


def maxchi(iphi):
    return lib().tg_asym_maxchi(iphi)


def symrho_plane(T, nslice, kmod, ith, Bz, Bxy, maxj):
    r, e, q = (np.zeros((361, 361)) for _ in range(3))
    info = np.zeros(5)
    rc = lib().tg_symrho_plane(T, nslice, kmod, ith, Bz, Bxy, maxj, _dp(r), _dp(e), _dp(q), _dp(info))
    if rc:
        raise RuntimeError("pmax too large")
    return r, e, q, info


def linden(T, nslice, bconst, npt, iodevn):
    out = np.zeros((npt, 4))
    info = np.zeros(4)
    lib().tg_linden(T, nslice, bconst, npt, iodevn, _dp(out), _dp(info))
    return out, info


def fmt_e15_8(v, scale1p=False):
    b = C.create_string_buffer(32)
    lib().tg_fmt_e15_8(float(v), 1 if scale1p else 0, b)
    return b.value.decode()


def rot_lines(out4):
    """the rows linden.f:68 writes: '(1p,7(1x,E15.8))'"""
    return ["".join(" " + fmt_e15_8(v, True) for v in row) + "\n" for row in out4]
