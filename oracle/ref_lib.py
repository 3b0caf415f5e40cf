"""ctypes binding of the UNMODIFIED reference library built by oracle/Makefile into
oracle/_ref/libtspws_ref.so (from /root/reference/src, never copied into this repo).

TEST INFRASTRUCTURE ONLY (see oracle/tspws_oracle.py header).  It exposes

* ``tspws_main``                       -- through the shared ctypes mirror of the ABI,
* the link-visible internals the reference never outputs: the frame
  (CreateWaveletFamily, src/FWTa/wavelet_def_v7.c:204), the analysis / synthesis
  (complex_1D_wavelet_dec / Re_complex_1D_wavelet_rec, src/FWTa/wavelet_v7.c:43,124) and
  the ST / PS stacks (tspws_stacks_float / partial_linear_stacks + tspws_stacks_double,
  src/ts_pws1f_lib.c:466,866,885) from which the coherence map is taken.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libtspws_ref.so")


class _WaveletFamily(C.Structure):          # src/FWTa/wavelet_v7.h:36-56
    _fields_ = [
        ("format", C.c_int), ("type", C.c_int), ("convtype", C.c_int),
        ("wframe", C.POINTER(C.POINTER(C.c_double))),
        ("wdualframe", C.POINTER(C.POINTER(C.c_double))),
        ("scale", C.POINTER(C.c_double)), ("Ls", C.POINTER(C.c_uint)), ("Lds", C.POINTER(C.c_uint)),
        ("center", C.POINTER(C.c_int)), ("center_df", C.POINTER(C.c_int)),
        ("Down_smp", C.POINTER(C.c_uint)), ("Ns", C.c_uint), ("V", C.c_uint),
        ("Cpsi", C.c_double), ("a0", C.c_double), ("b0", C.c_double), ("op1", C.c_double),
    ]


class _CWTvar(C.Structure):                 # src/FWTa/wavelet_v7.h:64-68
    _fields_ = [("d", C.POINTER(C.POINTER(C.c_double))), ("N", C.POINTER(C.c_uint)), ("S", C.c_uint)]


def available() -> bool:
    return os.path.exists(REF_SO)


_ref = None


def load():
    global _ref
    if _ref is not None:
        return _ref
    if not available():
        raise RuntimeError(f"{REF_SO} missing: run `make -C oracle` where /root/reference exists")
    lib = C.CDLL(REF_SO)
    lib.CreateWaveletFamily.restype = C.POINTER(_WaveletFamily)
    lib.CreateWaveletFamily.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_double, C.c_double,
                                        C.c_int, C.c_double, C.c_int]
    lib.DestroyWaveletFamily.argtypes = [C.POINTER(_WaveletFamily)]
    lib.CreateComplexWaveletVar.restype = C.POINTER(_CWTvar)
    lib.CreateComplexWaveletVar.argtypes = [C.POINTER(_WaveletFamily), C.c_uint]
    lib.DestroyComplexWaveletVar.argtypes = [C.POINTER(_CWTvar)]
    lib.complex_1D_wavelet_dec.argtypes = [C.POINTER(_CWTvar), C.POINTER(C.c_double), C.c_uint,
                                           C.POINTER(_WaveletFamily)]
    lib.Re_complex_1D_wavelet_rec.argtypes = [C.POINTER(C.c_double), C.POINTER(_CWTvar), C.c_uint,
                                              C.POINTER(_WaveletFamily)]
    lib.tspws_stacks_float.argtypes = [C.POINTER(_CWTvar), C.POINTER(_CWTvar), C.POINTER(C.c_float),
                                       C.c_void_p, C.c_int, C.c_uint, C.c_uint,
                                       C.POINTER(_WaveletFamily), C.POINTER(_CWTvar)]
    lib.tspws_stacks_double.argtypes = [C.POINTER(_CWTvar), C.POINTER(_CWTvar), C.POINTER(C.c_double),
                                        C.c_int, C.c_uint, C.c_uint,
                                        C.POINTER(_WaveletFamily), C.POINTER(_CWTvar)]
    lib.tspws_stacks_double.restype = None
    lib.partial_linear_stacks.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_void_p,
                                          C.c_uint, C.c_uint, C.c_uint, C.c_uint]
    lib.partial_linear_stacks.restype = None
    lib.SubsamplingPlan.argtypes = [C.c_char_p, C.c_uint, C.c_uint]
    lib.JackknifePlans.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_long), C.c_uint, C.c_uint,
                                   C.c_uint, C.c_uint]
    _ref = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class RefFrame:
    """A reference t_WaveletFamily plus numpy views of its geometry and filters."""

    def __init__(self, type_, J, V, N, s0, b0, w0, uni=0):
        lib = load()
        self.lib, self.N = lib, N
        self.pWF = lib.CreateWaveletFamily(type_, J, V, N, s0, b0, 0, w0, uni)
        if not self.pWF:
            raise RuntimeError("CreateWaveletFamily failed")
        f = self.pWF.contents
        S = f.Ns
        self.S, self.V, self.Cpsi = S, f.V, f.Cpsi
        self.scale = np.ctypeslib.as_array(f.scale, (S,)).copy()
        self.L = np.ctypeslib.as_array(f.Ls, (S,)).astype(np.int64)
        self.center = np.ctypeslib.as_array(f.center, (S,)).astype(np.int64)
        self.D = np.ctypeslib.as_array(f.Down_smp, (S,)).astype(np.int64)
        self.Ns = (N + self.D - 1) // self.D
        self.ncoef = int(self.Ns.sum())
        self.coef_off = np.concatenate([[0], np.cumsum(self.Ns)])
        self.w = []
        for s in range(S):
            a = np.ctypeslib.as_array(f.wframe[s], (2 * int(self.L[s]),)).copy()
            self.w.append(a[0::2] + 1j * a[1::2])

    def close(self):
        if self.pWF:
            self.lib.DestroyWaveletFamily(self.pWF)
            self.pWF = None

    def _var(self):
        v = self.lib.CreateComplexWaveletVar(self.pWF, self.N)
        if not v:
            raise MemoryError
        return v

    def _var_to_numpy(self, v):
        a = np.ctypeslib.as_array(v.contents.d[0], (2 * self.ncoef,)).copy()
        return a[0::2] + 1j * a[1::2]

    def analysis(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        v = self._var()
        self.lib.complex_1D_wavelet_dec(v, _dp(x), self.N, self.pWF)
        y = self._var_to_numpy(v)
        self.lib.DestroyComplexWaveletVar(v)
        return y

    def synthesis(self, Z):
        Z = np.ascontiguousarray(Z, dtype=np.complex128)
        v = self._var()
        C.memmove(v.contents.d[0], Z.ctypes.data, Z.nbytes)
        out = np.zeros(self.N)
        self.lib.Re_complex_1D_wavelet_rec(_dp(out), v, self.N, self.pWF)
        self.lib.DestroyComplexWaveletVar(v)
        return out

    def stacks(self, sigall, Kmax=0):
        """ST, PS, K of float32 traces [mtr, N] exactly as tspws_main computes them
        (src/ts_pws1f_lib.c:207-222): single stage if Kmax == 0 or Kmax > mtr."""
        x = np.ascontiguousarray(sigall, dtype=np.float32)
        mtr = x.shape[0]
        v, st, ps = self._var(), self._var(), self._var()
        if not Kmax or Kmax > mtr:
            K = mtr
            self.lib.tspws_stacks_float(st, ps, x.ctypes.data_as(C.POINTER(C.c_float)), None,
                                        self.N, self.N, mtr, self.pWF, v)
        else:
            K = Kmax
            g = np.zeros((K, self.N))
            self.lib.partial_linear_stacks(_dp(g), x.ctypes.data_as(C.POINTER(C.c_float)), None,
                                           self.N, self.N, K, mtr)
            self.lib.tspws_stacks_double(st, ps, _dp(g), self.N, self.N, K, self.pWF, v)
        ST, PS = self._var_to_numpy(st), self._var_to_numpy(ps)
        for a in (v, st, ps):
            self.lib.DestroyComplexWaveletVar(a)
        return ST, PS, K


def subsampling_plans(M, mtr, K):
    """M successive SubsamplingPlan() draws (src/ts_pws1f_lib.c:355-383) from the process's
    current rand() state -> sel[M, mtr] int8."""
    lib = load()
    sel = np.zeros((M, mtr), dtype=np.int8)
    for m in range(M):
        buf = C.create_string_buffer(mtr)
        lib.SubsamplingPlan(buf, mtr, K)
        sel[m] = np.frombuffer(buf.raw, dtype=np.int8, count=mtr)
    return sel


def jackknife_plans(times, n, d, Ccomb):
    lib = load()
    t = np.ascontiguousarray(times, dtype=np.int64)
    mtr = len(t)
    block = C.create_string_buffer(Ccomb * mtr)
    rows = (C.c_void_p * Ccomb)()
    base = C.addressof(block)
    for c in range(Ccomb):
        rows[c] = base + c * mtr
    lib.JackknifePlans.argtypes = [C.c_void_p, C.POINTER(C.c_long), C.c_uint, C.c_uint, C.c_uint, C.c_uint]
    rc = lib.JackknifePlans(C.cast(rows, C.c_void_p), t.ctypes.data_as(C.POINTER(C.c_long)), mtr, d, n, Ccomb)
    return rc, np.frombuffer(block.raw, dtype=np.int8).reshape(Ccomb, mtr).copy()
