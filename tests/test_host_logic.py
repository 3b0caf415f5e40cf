"""Host logic through the C-ABI (no GPU): parameter derivation, frame geometry and filters,
resampling plans -- against the numpy oracle (and the compiled reference when present)."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import tspws_oracle as O

CASES = [
    # (type, N, dt, extra t_tsPWS fields)
    (-1, 16384, 1.0, {}),
    (-1, 16384, 1.0, dict(w0=6.0, w0set=4)),
    (-3, 16384, 1.0, {}),
    (-1, 32768, 1.0, {}),
    (-1, 8192, 1.0, {}),
    (-1, 16501, 4.0, {}),
    (-1, 16501, 4.0, dict(fmin=0.004, J=3)),
    (-1, 16501, 4.0, dict(fmin=0.004)),
    (-2, 1000, 1.0, {}),
    (-1, 1000, 1.0, dict(Q=4.0, w0set=1)),
    (-1, 1000, 1.0, dict(cycle=3.0, w0set=2)),
    (-1, 1000, 1.0, dict(w0=3.0, w0set=4)),            # b0 truncates to 0 -> undecimated frame
    (-1, 1001, 1.0, dict(J=8, V=6, s0=4.0, lVfix=1, ls0fix=1)),
    (-3, 777, 1.0, dict(uni=1, J=4)),
    (-1, 1000, 1.0, dict(fmin=1e-9)),                  # fmin too low -> reset to default
]


@pytest.mark.parametrize("typ,N,dt,kw", CASES)
def test_params_and_frame_match_oracle(pkg, lib, typ, N, dt, kw):
    p = pkg.default_params(type=typ, **kw)
    assert lib.tspws_b200_derive_params(C.byref(p), N, dt) == 0
    po = O.derive_params(O.Params(type=typ, **kw), N, dt)
    assert (p.J, p.V) == (po.J, po.V)
    assert p.s0 == po.s0 and p.b0 == po.b0 and p.w0 == po.w0 and p.fmin == po.fmin

    fr = lib.tspws_b200_frame_create(p.type, p.J, p.V, N, p.s0, p.b0, p.w0, p.uni)
    assert fr
    try:
        d = pkg.frame_to_dict(fr)
    finally:
        lib.tspws_b200_frame_destroy(fr)
    fo = O.frame_from_params(po, N)
    assert d["S"] == fo.S and d["ncoef"] == fo.ncoef
    np.testing.assert_array_equal(d["L"], fo.L)
    np.testing.assert_array_equal(d["center"], fo.center)
    np.testing.assert_array_equal(d["D"], fo.D)
    np.testing.assert_array_equal(d["Ns"], fo.Ns)
    np.testing.assert_allclose(d["scale"], fo.scale, rtol=0, atol=0)
    assert abs(d["Cpsi"] - fo.Cpsi) <= 1e-15 * abs(fo.Cpsi)
    wo = np.concatenate(fo.w)
    assert np.max(np.abs(d["w"] - wo)) <= 2e-14 * np.max(np.abs(wo))


def test_geometry_of_baseline_configs(pkg, lib):
    """SURVEY.md Appendix B table (probed with the reference's CreateWaveletFamily)."""
    expect = {  # (type, N, w0) -> (V, J, scales, ncoef, ntaps)
        (-1, 16384, None): (4, 11, 44, 65504, 172907),
        (-1, 16384, 6.0): (5, 11, 55, 81880, 218177),
        (-3, 16384, None): (2, 10, 20, 98176, 24750),
        (-1, 32768, None): (4, 12, 48, 131040, 345853),
        (-1, 8192, None): (4, 10, 40, 32736, 86431),
    }
    for (typ, N, w0), (V, J, S, ncoef, ntaps) in expect.items():
        kw = dict(w0=w0, w0set=4) if w0 else {}
        p = pkg.default_params(type=typ, **kw)
        lib.tspws_b200_derive_params(C.byref(p), N, 1.0)
        assert (p.V, p.J) == (V, J)
        fr = lib.tspws_b200_frame_create(p.type, p.J, p.V, N, p.s0, p.b0, p.w0, p.uni)
        f = fr.contents
        assert (f.S, f.ncoef, f.ntaps) == (S, ncoef, ntaps)
        lib.tspws_b200_frame_destroy(fr)


def test_frame_rejects_bad_type(lib):
    assert not lib.tspws_b200_frame_create(0, 3, 4, 100, 2.0, 1.0, 5.0, 0)
    assert not lib.tspws_b200_frame_create(2, 3, 4, 100, 2.0, 1.0, 5.0, 0)
    assert not lib.tspws_b200_frame_create(-4, 3, 4, 100, 2.0, 1.0, 5.0, 0)


def test_subsampling_plan_counts(lib):
    for J, K in [(10, 3), (10, 5), (10, 8), (499, 250), (7, 7), (7, 0)]:
        buf = C.create_string_buffer(J)
        assert lib.tspws_b200_subsampling_plan(buf, J, K) == 0
        sel = np.frombuffer(buf.raw, dtype=np.int8, count=J)
        assert set(np.unique(sel)) <= {0, 1} and sel.sum() == K
    assert lib.tspws_b200_subsampling_plan(None, 4, 2) == 1
    assert lib.tspws_b200_subsampling_plan(C.create_string_buffer(4), 4, 5) == 2


def test_subsampling_plan_follows_reference_rand_sequence(lib):
    """Same rand() consumption as reference SubsamplingPlan: with both libraries in one process
    sharing libc's rand(), re-seeding and drawing must give identical plans."""
    from oracle import ref_lib as R
    if not R.available():
        pytest.skip("oracle/_ref not built")
    libc = C.CDLL(None)
    ref = R.load()
    for J, K in [(40, 10), (40, 30), (499, 250), (9, 4)]:
        a, b = C.create_string_buffer(J), C.create_string_buffer(J)
        libc.srand(1234)
        ref.SubsamplingPlan(a, J, K)
        libc.srand(1234)
        lib.tspws_b200_subsampling_plan(b, J, K)
        assert a.raw == b.raw


def test_jackknife_plans_match_oracle(lib):
    day = 86400
    t0 = 946684800  # 2000-01-01
    times = np.array([t0 + day * (3 * i) for i in range(200)], dtype=np.int64)
    for n, d in [(12, 2), (6, 1), (5, 3)]:
        Cn = math.comb(n, d)
        buf = C.create_string_buffer(Cn * len(times))
        rc = lib.tspws_b200_jackknife_plans(buf, times.ctypes.data_as(C.POINTER(C.c_long)), len(times), d, n, Cn)
        assert rc == 0
        sel = np.frombuffer(buf.raw, dtype=np.int8).reshape(Cn, len(times))
        np.testing.assert_array_equal(sel, O.jackknife_plans(times, n, d))
    assert lib.tspws_b200_jackknife_plans(None, None, 3, 1, 2, 2) == 1
    z = np.zeros(3, dtype=np.int64)
    assert lib.tspws_b200_jackknife_plans(C.create_string_buffer(6), z.ctypes.data_as(C.POINTER(C.c_long)), 3, 1, 2, 2) == -2


@pytest.mark.parametrize("kw", [dict(lrm=1), dict(type=-3, Kmax=2, unbiased=1, fold=1, wu=1.5),
                                dict(w0set=1, Q=4.0, subsmpl_N=3, subsmpl_p=0.5), dict(fmin=0.02, J=3, jackknife_n=4, jackknife_d=1, Kmax=2)])
def test_verbose_parameter_report_matches_reference_text(pkg, lib, capfd, kw):
    """`verbose` prints the derived parameters (src/ts_pws1f_lib.c:126-151); scripts grep that text.  Ours is
    printed before any CUDA call, so an empty stack (mtr = 0) shows it without a GPU; the reference prints the
    same block ahead of its stacking loops."""
    import ctypes as C
    import numpy as np
    from oracle import ref_lib as R
    if not R.available():
        pytest.skip("compiled reference missing")
    ref = pkg.declare_reference(R.load())
    libc = C.CDLL(None)
    n = 96
    x = np.random.default_rng(2).standard_normal((4, n)).astype(np.float32)

    def report(which, mtr):
        p = pkg.default_params(verbose=1, **kw)
        if mtr:
            pkg.call_tspws_main(which, x, p, times=np.arange(4, dtype=np.int64) * 86400 * 40 + 10**9)
        else:
            out, dat = pkg.TsPWSOut(), pkg.Data()
            dat.sigall = x.ctypes.data_as(C.POINTER(C.c_float))
            dat.hdr.max, dat.hdr.mtr, dat.hdr.dt, dat.hdr.beg = n, 0, 1.0, -(n - 1) / 2.0
            assert which.tspws_main(C.byref(p), C.byref(out), C.byref(dat)) == 0
        libc.fflush(None)
        text = capfd.readouterr().out
        a = text.index("Sequence length")
        return text[a:text.index("\n\n", a)]

    capfd.readouterr()
    assert report(lib, 0) == report(ref, 4)


@pytest.mark.parametrize("beg,expect_fold", [(-47.5, 1), (-47.0, 0), (0.0, 0), (-47.74, 1), (-47.76, 1), (-60.0, 1)])
def test_fold_rule_and_warning_match_reference(pkg, lib, capfd, beg, expect_fold):
    """fold is honoured only when 2*beg + (max-1)*dt <= dt/2 in float arithmetic, else a warning is printed and the
    flag cleared in the caller's block (src/ts_pws1f_lib.c:71-75) -- same decision, same text, same mutation."""
    import ctypes as C
    import numpy as np
    from oracle import ref_lib as R
    if not R.available():
        pytest.skip("compiled reference missing")
    ref = pkg.declare_reference(R.load())
    libc = C.CDLL(None)
    n = 96
    x = np.random.default_rng(3).standard_normal((3, n)).astype(np.float32)

    def run(which, mtr):
        p = pkg.default_params(fold=1)
        out, dat = pkg.TsPWSOut(), pkg.Data()
        xx = x.copy()
        ls, ts = np.zeros(n, np.float32), np.zeros(n, np.float32)
        dat.sigall = xx.ctypes.data_as(C.POINTER(C.c_float))
        dat.hdr.max, dat.hdr.mtr, dat.hdr.dt, dat.hdr.beg = n, mtr, 1.0, beg
        out.ls, out.tsPWS = ls.ctypes.data_as(C.POINTER(C.c_float)), ts.ctypes.data_as(C.POINTER(C.c_float))
        out.N, out.mtr = n, mtr
        assert which.tspws_main(C.byref(p), C.byref(out), C.byref(dat)) == 0
        libc.fflush(None)
        text = capfd.readouterr().out
        return p.fold, [l for l in text.splitlines() if l.startswith("Warning")]

    capfd.readouterr()
    ours, theirs = run(lib, 0), run(ref, 3)
    assert ours == theirs
    assert ours[0] == expect_fold and (len(ours[1]) == 0) == bool(expect_fold)


@pytest.mark.parametrize("typ", [0, -4, -7])
def test_unusable_wavelet_family_returns_4_like_the_reference(pkg, lib, capfd, typ):
    """CreateWaveletFamily fails -> tspws_main returns 4 (src/ts_pws1f_lib.c:199-204), whatever the input."""
    import ctypes as C
    import numpy as np
    from oracle import ref_lib as R
    n = 96
    x = np.random.default_rng(4).standard_normal((3, n)).astype(np.float32)

    def run(which, mtr):
        p = pkg.default_params(type=typ)
        out, dat = pkg.TsPWSOut(), pkg.Data()
        xx = x.copy()
        ls, ts = np.zeros(n, np.float32), np.zeros(n, np.float32)
        dat.sigall = xx.ctypes.data_as(C.POINTER(C.c_float))
        dat.hdr.max, dat.hdr.mtr, dat.hdr.dt, dat.hdr.beg = n, mtr, 1.0, -47.5
        out.ls, out.tsPWS = ls.ctypes.data_as(C.POINTER(C.c_float)), ts.ctypes.data_as(C.POINTER(C.c_float))
        out.N, out.mtr = n, mtr
        rc = which.tspws_main(C.byref(p), C.byref(out), C.byref(dat))
        C.CDLL(None).fflush(None)
        return rc, capfd.readouterr().out.splitlines()[0]

    capfd.readouterr()
    ours = [run(lib, 0), run(lib, 3)]                 # decided before CUDA is needed
    assert all(rc == 4 for rc, _ in ours)
    if R.available():
        theirs = run(pkg.declare_reference(R.load()), 3)
        assert theirs[0] == 4 and theirs[1] == ours[0][1]      # same first error line


def test_parameter_derivation_equals_the_compiled_reference_on_a_random_grid(pkg, lib):
    """src/ts_pws1f_lib.c:91-124 against our tspws_b200_derive_params, bit for bit, over 120 random parameter
    blocks (wavelet type, w0 / Q / cycles, fmin with and without J, fixed V / s0 / b0, uniform sampling): the
    reference mutates *tspws inside tspws_main, here run on a tiny two-trace input."""
    from oracle import ref_lib as R
    if not R.available():
        pytest.skip("compiled reference missing")
    ref = pkg.declare_reference(R.load())
    rng = np.random.default_rng(2024)
    checked = 0
    for _ in range(120):
        n = int(rng.choice([200, 256, 333, 500, 1000]))
        dt = float(rng.choice([1.0, 0.5, 4.0, 0.025]))
        kw = dict(type=int(rng.choice([-1, -1, -3, -2])))
        mode = int(rng.integers(0, 5))
        if kw["type"] != -3:
            if mode == 1:
                kw.update(w0set=1, Q=float(rng.uniform(2.5, 9.0)))
            elif mode == 2:
                kw.update(w0set=2, cycle=float(rng.uniform(1.5, 6.0)))
            elif mode == 3:
                kw.update(w0set=4, w0=float(rng.uniform(4.0, 12.0)))
        if rng.random() < 0.4:
            kw["fmin"] = float(rng.uniform(1.5, 12.0) / (dt * n))
            if rng.random() < 0.5:
                kw["J"] = int(rng.integers(1, 4))
        elif rng.random() < 0.3:
            kw["J"] = int(rng.integers(1, 4))
        if rng.random() < 0.2:
            kw.update(V=int(rng.integers(1, 7)), lVfix=1)
        if rng.random() < 0.2:
            kw.update(s0=float(rng.uniform(1.0, 4.0)), ls0fix=1)
        if rng.random() < 0.15:
            kw.update(b0=float(rng.choice([0.5, 1.0, 2.0])), lb0fix=1)
        if rng.random() < 0.15:
            kw["uni"] = 1
        ours = pkg.default_params(**kw)
        assert lib.tspws_b200_derive_params(C.byref(ours), n, dt) == 0
        if ours.J == 0 or ours.J > 7 or ours.V == 0:      # keep the reference's CPU run tiny and well defined
            continue
        x = rng.standard_normal((2, n)).astype(np.float32)
        r = pkg.call_tspws_main(ref, x, pkg.default_params(**kw), dt=dt)
        if r["rc"] != 0:
            continue
        theirs = r["params"]
        for f in ("J", "V", "s0", "b0", "w0", "fmin", "fold", "w0set"):
            assert getattr(ours, f) == getattr(theirs, f), (kw, n, dt, f, getattr(ours, f), getattr(theirs, f))
        checked += 1
    assert checked >= 60


def test_frame_equals_the_compiled_reference_on_a_random_grid(pkg, lib):
    """tspws_b200_frame_create against the reference's CreateWaveletFamily (src/FWTa/wavelet_def_v7.c:204-339)
    over 40 random frames: scales, lengths, centres, decimations bit-equal, filters and Cpsi to 1e-13."""
    from oracle import ref_lib as R
    if not R.available():
        pytest.skip("compiled reference missing")
    rng = np.random.default_rng(77)
    for _ in range(40):
        typ = int(rng.choice([-1, -2, -3]))
        N = int(rng.choice([257, 400, 512, 1000, 1501]))
        V = int(rng.integers(1, 6))
        J = int(rng.integers(1, 6))
        s0 = float(rng.choice([1.0, 2.0, 2.5, 4.0]))
        b0 = float(rng.choice([0.0, 0.5, 1.0, 2.0]))
        w0 = math.sqrt(2.0) if typ == -3 else float(rng.uniform(4.0, 9.0))
        uni = int(rng.random() < 0.2)
        rf = R.RefFrame(typ, J, V, N, s0, b0, w0, uni)
        fr = lib.tspws_b200_frame_create(typ, J, V, N, s0, b0, w0, uni)
        assert fr
        try:
            d = pkg.frame_to_dict(fr)
        finally:
            lib.tspws_b200_frame_destroy(fr)
        try:
            assert d["S"] == rf.S
            np.testing.assert_array_equal(d["L"], rf.L)
            np.testing.assert_array_equal(d["center"], rf.center)
            np.testing.assert_array_equal(d["D"], rf.D)
            np.testing.assert_array_equal(d["Ns"], rf.Ns)
            np.testing.assert_array_equal(d["scale"], rf.scale)
            assert abs(d["Cpsi"] - rf.Cpsi) <= 1e-13 * abs(rf.Cpsi)
            wr = np.concatenate(rf.w)
            assert np.max(np.abs(d["w"] - wr)) <= 1e-13 * np.max(np.abs(wr)), (typ, J, V, N, s0, b0, w0, uni)
        finally:
            rf.close()


@pytest.mark.parametrize("kw,N,dense_D,band_D", [
    (dict(w0=6.0, w0set=4), 16384, 256, 16),      # BASELINE configs[1] frame: octaves D = 16 .. 128, 5 voices
    (dict(), 16384, 256, 16),                     # default Morlet, 4 voices
    (dict(type=-3), 16384, 256, 16),              # Mexican hat, 2 voices (real filters)
    (dict(), 32768, 256, 32),                     # configs[3] length
    (dict(w0=6.0, w0set=4), 4096, 128, 8),        # short traces: few row blocks, octaves dropped when the window would wrap
], ids=["morlet_w6", "morlet_default", "mexhat", "n32768", "n4096"])
def test_banded_gemm_geometry_reproduces_the_transform(pkg, lib, kw, N, dense_D, band_D):
    """The banded tensor-core transform (csrc/tspws_mma.cu, fwd_dense_mma_kernel<true>) is an index rearrangement of
    src/FWTa/wavelet_v7.c:43-64: coefficient m = U P + u of a voice is sum_q row[P E + q] . B_q[u] over 64-sample
    row blocks.  This test takes the geometry from the C side (tspws_b200_band_layout = what the plan uploads),
    rebuilds the B tiles exactly as band_w_image_kernel does, multiplies in numpy and compares every covered
    coefficient with the oracle's analysis -- the layout logic is checked without a GPU."""
    p = pkg.default_params(**kw)
    assert lib.tspws_b200_derive_params(C.byref(p), N, 1.0) == 0
    fr = lib.tspws_b200_frame_create(p.type, p.J, p.V, N, p.s0, p.b0, p.w0, p.uni)
    assert fr
    try:
        rec = (C.c_int * (28 * 16))()
        s_first = C.c_int(-1)
        nb = lib.tspws_b200_band_layout(fr, dense_D, band_D, rec, 16, C.byref(s_first))
    finally:
        lib.tspws_b200_frame_destroy(fr)
    assert nb >= 1
    fo = O.frame_from_params(O.derive_params(O.Params(**kw), N, 1.0), N)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(N)
    Y = O.analysis(fo, x)
    off = fo.coef_off
    rows = x.reshape(N // 64, 64)
    kblocks = N // 64
    covered = np.zeros(fo.ncoef, dtype=bool)
    tiles = 0
    for i in range(nb):
        tile_base, nj, ncols, rpg, qmin, ngroups, nv, Pn, U, E, Ns, D = rec[28 * i:28 * i + 12]
        coef_rel = rec[28 * i + 12:28 * i + 20]
        s_of = rec[28 * i + 20:28 * i + 28]
        assert tile_base == tiles and ncols % 16 == 0 and nv * Pn * U * 2 <= ncols <= 128 and nj <= kblocks
        assert all(fo.D[s_of[v]] == D for v in range(nv)) and Ns == fo.Ns[s_of[0]]
        tiles += nj
        B = np.zeros((nj, ncols, 64))
        for col in range(ncols):
            t2 = col >> 1; u = t2 % U; t3 = t2 // U; pp = t3 % Pn; vi = t3 // Pn
            if vi >= nv:
                continue
            s = s_of[vi]
            w, L, c = fo.w[s], int(fo.L[s]), int(fo.center[s])
            for jj in range(nj):
                l = 64 * (jj + qmin - E * pp) + np.arange(64) - u * D + c
                ok = (l >= 0) & (l < L)
                v = (-w.imag if col & 1 else w.real)[np.clip(l, 0, L - 1)]
                B[jj, col] = np.where(ok, v, 0.0)
        for g in range(ngroups):
            row0 = (g * rpg + qmin) % kblocks
            acc = np.zeros(ncols)
            for jj in range(nj):
                acc += B[jj] @ rows[(row0 + jj) % kblocks]
            for col in range(ncols):
                t2 = col >> 1; u = t2 % U; t3 = t2 // U; pp = t3 % Pn; vi = t3 // Pn
                m = (g * Pn + pp) * U + u
                if vi < nv and m < Ns:
                    ci = off[s_first.value] + coef_rel[vi] + m
                    assert ci == off[s_of[vi]] + m
                    ref = Y[ci].imag if col & 1 else Y[ci].real
                    assert abs(acc[col] - ref) <= 1e-11 * (1 + np.abs(Y).max())
                    if not (col & 1):
                        covered[ci] = True
    # every coefficient of the covered octaves is produced exactly once, nothing outside them
    lo, hi = off[s_first.value], off[max(rec[20:20 + rec[6]]) + 1]                 # record 0 is the coarsest octave
    assert covered[lo:hi].all() and not covered[:lo].any() and not covered[hi:].any()
