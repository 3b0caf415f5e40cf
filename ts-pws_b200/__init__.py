"""ts-pws_b200 -- thin ctypes mirror of the C ABI in include/tspws_b200.h.

The product is libtspws_b200.so (C host code + sm_100a CUDA kernels, built by
``make -C ts-pws_b200/csrc``).  This module only mirrors the reference's
parameter blocks (reference src/ts_pws1f_lib.h:19-102) so that tests and
bench.py can call ``tspws_main`` with host buffers exactly like the reference's
CLI (src/ts_pws1f.c:307) or MEX gateway (src/gw_ts_pws.c:164) do.  The very
same mirror is used to drive the compiled reference (oracle/_ref) -- the two
libraries share the ABI, which is the point of the drop-in boundary.

There is no CPU fallback here: if the shared library is missing, loading fails
loudly.
"""
from __future__ import annotations

import ctypes as C
import math
import time
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libtspws_b200.so")

TSPWS_ERR_CUDA = 77
OPT_WRITEBACK, OPT_FP64_TAIL, OPT_DEVICE, OPT_SS_JACKKNIFE = 1, 2, 3, 4

W0_DEFAULT = math.pi * math.sqrt(2.0 / math.log(2.0))


class TsPWS(C.Structure):
    """t_tsPWS (reference src/ts_pws1f_lib.h:19-57)."""
    _fields_ = [
        ("type", C.c_int), ("uni", C.c_uint), ("J", C.c_uint), ("V", C.c_uint),
        ("s0", C.c_double), ("b0", C.c_double), ("w0", C.c_double), ("wu", C.c_double),
        ("fmin", C.c_double), ("Q", C.c_double), ("cycle", C.c_double),
        ("w0set", C.c_int), ("lrm", C.c_int), ("bin", C.c_int), ("lkinst", C.c_int),
        ("lVfix", C.c_int), ("ls0fix", C.c_int), ("lb0fix", C.c_int), ("verbose", C.c_int),
        ("fold", C.c_int), ("unbiased", C.c_int), ("convergence", C.c_int),
        ("subsmpl_N", C.c_uint), ("subsmpl_p", C.c_double),
        ("jackknife_n", C.c_uint), ("jackknife_d", C.c_uint), ("obin", C.c_uint),
        ("AllSteps", C.c_int), ("Nmax", C.c_uint), ("Kmax", C.c_uint),
        ("kinst", C.c_char_p), ("filein", C.c_char_p), ("fileout", C.c_char_p),
        ("fileconv", C.c_char_p),
    ]


class Hdr(C.Structure):
    """t_hdr (reference src/ts_pws1f_lib.h:59-77)."""
    _fields_ = [
        ("max", C.c_int), ("mtr", C.c_uint),
        ("evla", C.c_float), ("evlo", C.c_float), ("stla", C.c_float), ("stlo", C.c_float),
        ("stel", C.c_float), ("dt", C.c_float), ("beg", C.c_float),
        ("net1", C.c_char * 9), ("sta1", C.c_char * 9), ("loc1", C.c_char * 9), ("chn1", C.c_char * 9),
        ("net2", C.c_char * 9), ("sta2", C.c_char * 9), ("loc2", C.c_char * 9), ("chn2", C.c_char * 9),
    ]


class TsPWSOut(C.Structure):
    """t_tsPWS_out (reference src/ts_pws1f_lib.h:79-94)."""
    _fields_ = [
        ("ls", C.POINTER(C.c_float)), ("tsPWS", C.POINTER(C.c_float)),
        ("ls_sim", C.POINTER(C.c_double)), ("tsPWS_sim", C.POINTER(C.c_double)),
        ("ls_misfit", C.POINTER(C.c_double)), ("tsPWS_misfit", C.POINTER(C.c_double)),
        ("ls_steps", C.POINTER(C.c_float)), ("tsPWS_steps", C.POINTER(C.c_float)),
        ("ls_subsmpl", C.POINTER(C.POINTER(C.c_float))),
        ("tsPWS_subsmpl", C.POINTER(C.POINTER(C.c_float))),
        ("mtr_subsmpl", C.POINTER(C.c_uint)),
        ("M", C.c_uint), ("N", C.c_uint), ("mtr", C.c_uint),
    ]


class Data(C.Structure):
    """t_data (reference src/ts_pws1f_lib.h:96-102)."""
    _fields_ = [
        ("sigall", C.POINTER(C.c_float)), ("time", C.POINTER(C.c_long)),
        ("lag0", C.POINTER(C.c_float)), ("hdr", Hdr), ("reference", C.POINTER(C.c_float)),
    ]


assert C.sizeof(TsPWS) == 184 and C.sizeof(Hdr) == 108
assert C.sizeof(TsPWSOut) == 104 and C.sizeof(Data) == 144


class Frame(C.Structure):
    """tspws_frame_t (include/tspws_b200.h)."""
    _fields_ = [
        ("type", C.c_int), ("S", C.c_uint), ("N", C.c_uint), ("V", C.c_uint),
        ("Cpsi", C.c_double), ("ncoef", C.c_size_t), ("ntaps", C.c_size_t),
        ("scale", C.POINTER(C.c_double)), ("L", C.POINTER(C.c_uint)),
        ("center", C.POINTER(C.c_int)), ("D", C.POINTER(C.c_uint)),
        ("Ns", C.POINTER(C.c_uint)), ("coef_off", C.POINTER(C.c_size_t)),
        ("tap_off", C.POINTER(C.c_size_t)), ("w", C.POINTER(C.c_double)),
    ]


def default_params(**kw) -> TsPWS:
    """The CLI's defaults (reference src/ts_pws1f.c:140-142): Morlet, V=4, s0=2, b0=1,
    w0=pi*sqrt(2/ln2), nu=2, cycle=2, everything else 0 (J=0 -> derived)."""
    p = TsPWS()
    p.type, p.V, p.s0, p.b0, p.w0, p.wu, p.cycle = -1, 4, 2.0, 1.0, W0_DEFAULT, 2.0, 2.0
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(f"t_tsPWS has no field {k!r}")
        setattr(p, k, v)
    return p


_lib: Optional[C.CDLL] = None


def _declare(lib: C.CDLL, full: bool) -> C.CDLL:
    lib.tspws_main.restype = C.c_int
    lib.tspws_main.argtypes = [C.POINTER(TsPWS), C.POINTER(TsPWSOut), C.POINTER(Data)]
    if not full:
        return lib
    lib.tspws_b200_main_batch.restype = C.c_int
    lib.tspws_b200_main_batch.argtypes = [C.POINTER(TsPWS), C.POINTER(TsPWSOut), C.POINTER(Data), C.c_uint]
    lib.tspws_b200_derive_params.restype = C.c_int
    lib.tspws_b200_derive_params.argtypes = [C.POINTER(TsPWS), C.c_uint, C.c_float]
    lib.tspws_b200_band_layout.restype = C.c_int
    lib.tspws_b200_band_layout.argtypes = [C.POINTER(Frame), C.c_long, C.c_long, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
    lib.tspws_b200_frame_create.restype = C.POINTER(Frame)
    lib.tspws_b200_frame_create.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint,
                                            C.c_double, C.c_double, C.c_double, C.c_int]
    lib.tspws_b200_frame_destroy.restype = None
    lib.tspws_b200_frame_destroy.argtypes = [C.POINTER(Frame)]
    lib.tspws_b200_subsampling_plan.restype = C.c_int
    lib.tspws_b200_subsampling_plan.argtypes = [C.c_char_p, C.c_uint, C.c_uint]
    lib.tspws_b200_jackknife_plans.restype = C.c_int
    lib.tspws_b200_jackknife_plans.argtypes = [C.c_char_p, C.POINTER(C.c_long), C.c_uint,
                                               C.c_uint, C.c_uint, C.c_uint]
    lib.tspws_b200_get_coherence.restype = C.c_long
    lib.tspws_b200_get_coherence.argtypes = [C.POINTER(C.c_double), C.c_size_t]
    lib.tspws_b200_get_stacks.restype = C.c_long
    lib.tspws_b200_get_stacks.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_size_t]
    lib.tspws_b200_set_selection.restype = C.c_int
    lib.tspws_b200_set_selection.argtypes = [C.c_char_p, C.c_uint, C.c_uint]
    lib.tspws_b200_set_option.restype = C.c_int
    lib.tspws_b200_set_option.argtypes = [C.c_int, C.c_long]
    lib.tspws_b200_get_option.restype = C.c_long
    lib.tspws_b200_get_option.argtypes = [C.c_int]
    lib.tspws_b200_device_count.restype = C.c_int
    lib.tspws_b200_shutdown.restype = None
    lib.tspws_b200_version.restype = C.c_char_p
    lib.tspws_b200_kernel_launches.restype = C.c_ulong
    lib.tspws_b200_profile.argtypes = [C.c_int]
    lib.tspws_b200_profile_read.argtypes = [C.c_char_p, C.c_int]
    lib.tspws_b200_load_sac_list.argtypes = [C.c_char_p, C.POINTER(Data), C.c_int]
    lib.tspws_b200_load_bin.argtypes = [C.c_char_p, C.POINTER(Data), C.c_int]
    lib.tspws_b200_save_bin.argtypes = [C.c_char_p, C.POINTER(Data), C.c_char_p]
    lib.tspws_b200_write_sac.argtypes = [C.c_char_p, C.POINTER(C.c_float), C.POINTER(Hdr), C.c_int, C.c_float, C.c_char_p, C.c_float]
    lib.tspws_b200_read_sac.restype = C.c_long
    lib.tspws_b200_read_sac.argtypes = [C.c_char_p, C.POINTER(C.c_float), C.c_long, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]
    lib.tspws_b200_free_data.restype = None
    lib.tspws_b200_free_data.argtypes = [C.POINTER(Data)]
    lib.tspws_b200_plan_create.restype = C.c_void_p
    lib.tspws_b200_plan_create.argtypes = [C.POINTER(TsPWS), C.c_int, C.c_uint, C.c_uint, C.c_uint]
    lib.tspws_b200_plan_destroy.restype = None
    lib.tspws_b200_plan_destroy.argtypes = [C.c_void_p]
    lib.tspws_b200_plan_reduce.restype = C.c_int
    lib.tspws_b200_plan_reduce.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tspws_b200_plan_exchange_ptr.restype = C.c_void_p
    lib.tspws_b200_plan_exchange_ptr.argtypes = [C.c_void_p]
    lib.tspws_b200_plan_exchange_bytes.restype = C.c_size_t
    lib.tspws_b200_plan_exchange_bytes.argtypes = [C.c_void_p]
    lib.tspws_b200_plan_exchange_is_f64.restype = C.c_int
    lib.tspws_b200_plan_exchange_is_f64.argtypes = [C.c_void_p]
    lib.tspws_b200_plan_bind_exchange.restype = C.c_int
    lib.tspws_b200_plan_bind_exchange.argtypes = [C.c_void_p, C.c_void_p]
    lib.tspws_b200_plan_mark_exchanged.restype = C.c_int
    lib.tspws_b200_plan_mark_exchanged.argtypes = [C.c_void_p]
    lib.tspws_b200_plan_peer_export.restype = C.c_int
    lib.tspws_b200_plan_peer_export.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.tspws_b200_plan_peer_import.restype = C.c_int
    lib.tspws_b200_plan_peer_import.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.tspws_b200_plan_peer_region.restype = C.c_void_p
    lib.tspws_b200_plan_peer_region.argtypes = [C.c_void_p, C.c_int]
    lib.tspws_b200_plan_peer_import_ptrs.restype = C.c_int
    lib.tspws_b200_plan_peer_import_ptrs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.tspws_b200_plan_peer_error.restype = C.c_int
    lib.tspws_b200_plan_peer_error.argtypes = [C.c_void_p]
    lib.tspws_b200_plan_peer_close.restype = C.c_int
    lib.tspws_b200_plan_peer_close.argtypes = [C.c_void_p]
    lib.tspws_b200_plan_finish.restype = C.c_int
    lib.tspws_b200_plan_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tspws_b200_plan_part_align.restype = C.c_uint
    lib.tspws_b200_plan_part_align.argtypes = [C.c_void_p]
    lib.tspws_b200_plan_reduce_part.restype = C.c_int
    lib.tspws_b200_plan_reduce_part.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_void_p]
    lib.tspws_b200_plan_reduce_done.restype = C.c_int
    lib.tspws_b200_plan_reduce_done.argtypes = [C.c_void_p, C.c_void_p]
    lib.tspws_b200_plan_stack.restype = C.c_int
    lib.tspws_b200_plan_stack.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tspws_b200_pipe_create.restype = C.c_void_p
    lib.tspws_b200_pipe_create.argtypes = [C.POINTER(TsPWS), C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_int]
    lib.tspws_b200_pipe_submit.restype = C.c_int
    lib.tspws_b200_pipe_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tspws_b200_pipe_sync.restype = C.c_int
    lib.tspws_b200_pipe_sync.argtypes = [C.c_void_p]
    lib.tspws_b200_pipe_stream.restype = C.c_void_p
    lib.tspws_b200_pipe_stream.argtypes = [C.c_void_p, C.c_int]
    lib.tspws_b200_pipe_plan.restype = C.c_void_p
    lib.tspws_b200_pipe_plan.argtypes = [C.c_void_p, C.c_int]
    lib.tspws_b200_pipe_mode.restype = C.c_int
    lib.tspws_b200_pipe_mode.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.tspws_b200_pipe_destroy.restype = None
    lib.tspws_b200_pipe_destroy.argtypes = [C.c_void_p]
    lib.tspws_b200_plan_stacks_ptr.restype = C.c_int
    lib.tspws_b200_plan_stacks_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p),
                                               C.POINTER(C.c_void_p), C.POINTER(C.c_uint),
                                               C.POINTER(C.c_int)]
    lib.tspws_b200_plan_get_stacks.restype = C.c_long
    lib.tspws_b200_plan_get_stacks.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_size_t]
    lib.tspws_b200_plan_K.restype = C.c_uint
    lib.tspws_b200_plan_K.argtypes = [C.c_void_p]
    lib.tspws_b200_plan_ncoef.restype = C.c_size_t
    lib.tspws_b200_plan_ncoef.argtypes = [C.c_void_p]
    return lib


EXPORTS = [
    "tspws_main", "tspws_b200_main_batch", "tspws_b200_derive_params", "tspws_b200_frame_create", "tspws_b200_frame_destroy", "tspws_b200_band_layout",
    "tspws_b200_subsampling_plan", "tspws_b200_jackknife_plans", "tspws_b200_get_coherence",
    "tspws_b200_get_stacks", "tspws_b200_set_selection", "tspws_b200_set_option",
    "tspws_b200_get_option", "tspws_b200_device_count", "tspws_b200_shutdown", "tspws_b200_version",
    "tspws_b200_kernel_launches", "tspws_b200_profile", "tspws_b200_profile_reset", "tspws_b200_profile_read", "tspws_b200_plan_create", "tspws_b200_plan_destroy",
    "tspws_b200_plan_reduce", "tspws_b200_plan_exchange_ptr", "tspws_b200_plan_exchange_bytes",
    "tspws_b200_load_sac_list", "tspws_b200_load_bin", "tspws_b200_save_bin", "tspws_b200_save_rows_bin", "tspws_b200_write_sac",
    "tspws_b200_read_sac", "tspws_b200_free_data",
    "tspws_b200_plan_exchange_is_f64", "tspws_b200_plan_bind_exchange", "tspws_b200_plan_mark_exchanged", "tspws_b200_plan_peer_export", "tspws_b200_plan_peer_import", "tspws_b200_plan_peer_region", "tspws_b200_plan_peer_import_ptrs", "tspws_b200_plan_peer_error", "tspws_b200_plan_peer_close", "tspws_b200_plan_part_align", "tspws_b200_plan_reduce_part", "tspws_b200_plan_reduce_done", "tspws_b200_plan_finish", "tspws_b200_plan_stack",
    "tspws_b200_plan_stacks_ptr", "tspws_b200_plan_ncoef", "tspws_b200_plan_get_stacks", "tspws_b200_plan_K",
    "tspws_b200_pipe_create", "tspws_b200_pipe_submit", "tspws_b200_pipe_sync", "tspws_b200_pipe_stream", "tspws_b200_pipe_plan", "tspws_b200_pipe_mode", "tspws_b200_pipe_destroy",
]


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Load libtspws_b200.so (built in-tree).  Raises if it is missing -- no fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build it with `make -C ts-pws_b200/csrc` "
            "(or __graft_entry__.build()). There is no CPU fallback.")
    lib = _declare(C.CDLL(p, mode=C.RTLD_GLOBAL), full=True)
    if path is None:
        _lib = lib
    return lib


def declare_reference(lib: C.CDLL) -> C.CDLL:
    """Give a compiled reference library (oracle/_ref) the same tspws_main signature."""
    return _declare(lib, full=False)


# ----------------------------------------------------------------------------------------
# Host-buffer convenience wrapper around tspws_main (any ABI-compatible library)
# ----------------------------------------------------------------------------------------
def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def call_tspws_main(lib: C.CDLL, sigall: np.ndarray, params: TsPWS, dt: float = 1.0,
                    beg: Optional[float] = None, times: Optional[np.ndarray] = None,
                    reference: Optional[np.ndarray] = None, copy: bool = True) -> dict:
    """Run ``tspws_main`` of ``lib`` on ``sigall`` ([mtr, max] float32).

    Allocates the outputs exactly like the CLI does (reference src/ts_pws1f.c:229-268,
    including M for resampling) and returns them in a dict together with the mutated
    parameter block.  ``sigall`` is copied first unless copy=False (the call mutates it).
    """
    x = np.array(sigall, dtype=np.float32, order="C", copy=copy)
    if x.ndim != 2:
        raise ValueError("sigall must be [mtr, max]")
    mtr_all, nmax = x.shape
    p = TsPWS.from_buffer_copy(params)
    out, dat = TsPWSOut(), Data()
    dat.sigall = _fptr(x)
    dat.hdr.max, dat.hdr.mtr, dat.hdr.dt = nmax, mtr_all, dt
    dat.hdr.beg = -(nmax - 1) * dt / 2.0 if beg is None else beg
    keep = [x]
    if times is not None:
        t = np.ascontiguousarray(times, dtype=np.int64)
        dat.time = t.ctypes.data_as(C.POINTER(C.c_long))
        keep.append(t)
    if reference is not None:
        r = np.ascontiguousarray(reference, dtype=np.float32)
        dat.reference = _fptr(r)
        keep.append(r)
    mtr = p.Nmax if p.Nmax else mtr_all
    res = {}
    ls = np.zeros(nmax, np.float32)
    ts = np.zeros(nmax, np.float32)
    out.ls, out.tsPWS, out.N, out.mtr = _fptr(ls), _fptr(ts), nmax, mtr
    if p.convergence:
        for name in ("ls_sim", "tsPWS_sim", "ls_misfit", "tsPWS_misfit"):
            a = np.zeros(mtr_all, np.float64)
            setattr(out, name, a.ctypes.data_as(C.POINTER(C.c_double)))
            res[name] = a
        if p.AllSteps:
            for name in ("ls_steps", "tsPWS_steps"):
                a = np.zeros((mtr_all, nmax), np.float32)
                setattr(out, name, _fptr(a))
                res[name] = a
    # resampling bookkeeping as in the CLI (src/ts_pws1f.c:244-277)
    M = 0
    if p.jackknife_n:
        p.subsmpl_N, p.subsmpl_p = 0, 0.0
        if p.jackknife_d == 0 or p.jackknife_d >= p.jackknife_n:
            p.jackknife_d = p.jackknife_n = 0
        else:
            M = math.comb(p.jackknife_n, p.jackknife_d)
    if p.subsmpl_N:
        if p.subsmpl_p < 0 or p.subsmpl_p > 1:
            p.subsmpl_N, p.subsmpl_p = 0, 0.0
        else:
            M = p.subsmpl_N
    if M:
        out.M = M
        msub = np.zeros(M, np.uint32)
        lsub = np.zeros((M, nmax), np.float32)
        tsub = np.zeros((M, nmax), np.float32)
        lsub.fill(0); tsub.fill(0)          # touch the pages now: first-touch faults are not the library's time
        rows_l = (C.POINTER(C.c_float) * M)(*[_fptr(lsub[m]) for m in range(M)])
        rows_t = (C.POINTER(C.c_float) * M)(*[_fptr(tsub[m]) for m in range(M)])
        out.mtr_subsmpl = msub.ctypes.data_as(C.POINTER(C.c_uint))
        out.ls_subsmpl = C.cast(rows_l, C.POINTER(C.POINTER(C.c_float)))
        out.tsPWS_subsmpl = C.cast(rows_t, C.POINTER(C.POINTER(C.c_float)))
        keep += [rows_l, rows_t]
        res.update(mtr_subsmpl=msub, ls_subsmpl=lsub, tsPWS_subsmpl=tsub)
    t0 = time.perf_counter()
    rc = lib.tspws_main(C.byref(p), C.byref(out), C.byref(dat))
    res.update(rc=rc, ls=ls, tsPWS=ts, params=p, sigall=x, M=M, seconds=time.perf_counter() - t0)
    return res


def call_main_batch(lib: C.CDLL, pairs, params: TsPWS, dt: float = 1.0, out_ls=None, out_ts=None) -> dict:
    """``tspws_b200_main_batch`` on a sequence of [mtr_i, max_i] float32 arrays (one per station pair).
    The arrays are used in place when already float32 C-contiguous (pinned torch tensors' numpy views
    work).  Returns {'rc', 'ls': [..], 'tsPWS': [..], 'params'}."""
    n = len(pairs)
    p = TsPWS.from_buffer_copy(params)
    outs, ins = (TsPWSOut * n)(), (Data * n)()
    xs, ls, ts = [], [], []
    for i, a in enumerate(pairs):
        x = np.ascontiguousarray(a, dtype=np.float32)
        mtr, nmax = x.shape
        l = np.zeros(nmax, np.float32) if out_ls is None else out_ls[i]
        t = np.zeros(nmax, np.float32) if out_ts is None else out_ts[i]
        ins[i].sigall = _fptr(x)
        ins[i].hdr.max, ins[i].hdr.mtr, ins[i].hdr.dt, ins[i].hdr.beg = nmax, mtr, dt, -(nmax - 1) * dt / 2.0
        outs[i].ls, outs[i].tsPWS, outs[i].N, outs[i].mtr = _fptr(l), _fptr(t), nmax, mtr
        xs.append(x); ls.append(l); ts.append(t)
    rc = lib.tspws_b200_main_batch(C.byref(p), outs, ins, n)
    return {"rc": rc, "ls": ls, "tsPWS": ts, "params": p, "_keep": xs}


def read_profile(lib) -> dict:
    """Per-kernel in-stream timings collected since the last reset: name -> (launches, total_ms)."""
    buf = C.create_string_buffer(8192)
    lib.tspws_b200_profile_read(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        out[name] = (int(cnt), float(ms))
    return out


def frame_to_dict(fr) -> dict:
    """Copy a tspws_frame_t* into numpy arrays."""
    f = fr.contents
    S = f.S
    d = dict(type=f.type, S=S, N=f.N, V=f.V, Cpsi=f.Cpsi, ncoef=f.ncoef, ntaps=f.ntaps)
    d["scale"] = np.ctypeslib.as_array(f.scale, (S,)).copy()
    d["L"] = np.ctypeslib.as_array(f.L, (S,)).copy()
    d["center"] = np.ctypeslib.as_array(f.center, (S,)).copy()
    d["D"] = np.ctypeslib.as_array(f.D, (S,)).copy()
    d["Ns"] = np.ctypeslib.as_array(f.Ns, (S,)).copy()
    d["coef_off"] = np.ctypeslib.as_array(f.coef_off, (S + 1,)).copy()
    d["tap_off"] = np.ctypeslib.as_array(f.tap_off, (S + 1,)).copy()
    w = np.ctypeslib.as_array(f.w, (2 * f.ntaps,)).copy()
    d["w"] = w[0::2] + 1j * w[1::2]
    return d
