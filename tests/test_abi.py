"""The C-ABI library loads without a GPU and exports every symbol include/tspws_b200.h declares;
struct layouts match the reference ABI."""
import ctypes as C
import os
import re

from conftest import ROOT


def test_exports_match_header(pkg, lib):
    hdr = open(os.path.join(ROOT, "include", "tspws_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(tspws_main|tspws_b200_\w+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(pkg.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_struct_layout(pkg):
    T = pkg.TsPWS
    assert C.sizeof(T) == 184
    off = {f[0]: getattr(T, f[0]).offset for f in T._fields_}
    assert (off["type"], off["uni"], off["J"], off["V"]) == (0, 4, 8, 12)
    assert (off["s0"], off["b0"], off["w0"], off["wu"], off["fmin"], off["Q"], off["cycle"]) == (16, 24, 32, 40, 48, 56, 64)
    assert (off["w0set"], off["lrm"], off["verbose"], off["fold"], off["unbiased"], off["convergence"]) == (72, 76, 100, 104, 108, 112)
    assert (off["subsmpl_N"], off["subsmpl_p"], off["jackknife_n"], off["jackknife_d"], off["obin"]) == (116, 120, 128, 132, 136)
    assert (off["AllSteps"], off["Nmax"], off["Kmax"], off["kinst"], off["fileconv"]) == (140, 144, 148, 152, 176)
    assert C.sizeof(pkg.Hdr) == 108
    assert C.sizeof(pkg.TsPWSOut) == 104 and pkg.TsPWSOut.M.offset == 88 and pkg.TsPWSOut.mtr.offset == 96
    assert C.sizeof(pkg.Data) == 144 and pkg.Data.hdr.offset == 24 and pkg.Data.reference.offset == 136


def test_null_arguments(pkg, lib):
    assert lib.tspws_main(None, None, None) == -1
    p = pkg.default_params()
    assert lib.tspws_main(C.byref(p), None, None) == -1


def test_null_arguments_of_the_additive_entry_points(pkg, lib):
    """Argument checks that come before any CUDA call (so they run without a GPU)."""
    p = pkg.default_params()
    assert lib.tspws_b200_main_batch(None, None, None, 3) == -1
    assert lib.tspws_b200_main_batch(C.byref(p), None, None, 0) == -1
    h = (C.c_ubyte * 64)()
    assert lib.tspws_b200_plan_peer_export(None, 2, h) != 0
    assert lib.tspws_b200_plan_peer_import(None, 0, 2, h) != 0
    assert lib.tspws_b200_plan_peer_error(None) == 0
    assert lib.tspws_b200_plan_reduce(None, None, None) != 0
    assert lib.tspws_b200_plan_finish(None, None, None, None) != 0
    assert lib.tspws_b200_plan_exchange_bytes(None) == 0
    assert not lib.tspws_b200_plan_create(None, 16, 1, 1, 0)
    assert not lib.tspws_b200_plan_create(C.byref(p), 0, 1, 1, 0)          # max <= 0
    assert not lib.tspws_b200_plan_create(C.byref(p), 16, 2, 1, 0)         # more local than total traces


def test_version_and_options(pkg, lib):
    assert b"tspws_b200" in lib.tspws_b200_version()
    assert lib.tspws_b200_set_option(pkg.OPT_WRITEBACK, 1) == 0
    assert lib.tspws_b200_get_option(pkg.OPT_WRITEBACK) == 1
    assert lib.tspws_b200_set_option(pkg.OPT_WRITEBACK, 0) == 0
    assert lib.tspws_b200_set_option(99, 0) != 0
    assert lib.tspws_b200_set_option(pkg.OPT_SS_JACKKNIFE, 1) == 0 and lib.tspws_b200_get_option(pkg.OPT_SS_JACKKNIFE) == 1
    assert lib.tspws_b200_set_option(pkg.OPT_SS_JACKKNIFE, 0) == 0


def test_empty_input_is_a_no_op_before_any_cuda_call(pkg, lib):
    """No traces (or no samples): the reference falls through its loops and returns 0
    (src/ts_pws1f_lib.c:194); here that happens before CUDA is touched, so it also holds without a GPU.
    The parameter block is still derived (mutated) like in the reference."""
    import numpy as np
    p = pkg.default_params(lrm=1)
    out, dat = pkg.TsPWSOut(), pkg.Data()
    x = np.zeros((1, 64), np.float32)
    ls = np.full(64, 7.0, np.float32)
    ts = np.full(64, 7.0, np.float32)
    dat.sigall = x.ctypes.data_as(C.POINTER(C.c_float))
    dat.hdr.max, dat.hdr.mtr, dat.hdr.dt, dat.hdr.beg = 64, 0, 1.0, -31.5
    out.ls = ls.ctypes.data_as(C.POINTER(C.c_float))
    out.tsPWS = ts.ctypes.data_as(C.POINTER(C.c_float))
    assert lib.tspws_main(C.byref(p), C.byref(out), C.byref(dat)) == 0
    assert (ls == 7.0).all() and (ts == 7.0).all()
    assert p.J > 0 and p.V == 4                      # derived from nsamp = 64 although nothing was stacked
    dat.hdr.max, dat.hdr.mtr = 0, 5
    assert lib.tspws_main(C.byref(p), C.byref(out), C.byref(dat)) == 0
