"""BASELINE.json configs[3] and configs[4] at their named sizes (C4, C5), on the GPU,
through the C-ABI: resampling over 2000 x 32768 with M = 1000 realisations, and a batch of independent
station pairs of 500 x 8192.  Checked against the numpy oracle on the realisations / pairs the oracle
can do in seconds and through size-independent identities on the rest.
Tolerance: stacked traces <= 1e-4 (max|a-b| / max|b|)."""
import ctypes as C

import numpy as np
import pytest

from oracle import cases as CS, tspws_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


# ---------------------------------------------------------------------------------- batch (C5)
def _pairs(npairs, mtr, n, seed0):
    return [O.synthetic_traces(mtr, n, seed=seed0 + 7 * i, amp=0.3 + 0.05 * (i % 5)) for i in range(npairs)]


@pytest.mark.parametrize("kw", [dict(lrm=1), dict(lrm=1, Kmax=6, unbiased=1), dict(fold=1, lrm=1, type=-3)])
def test_batch_equals_call_by_call(pkg, gpu_lib, kw):
    """main_batch == tspws_main in a loop, bit for bit (same kernels, same order), and == oracle."""
    pairs = _pairs(5, 40, 1024, 100)
    b = pkg.call_main_batch(gpu_lib, pairs, pkg.default_params(**kw))
    assert b["rc"] == 0
    for i, x in enumerate(pairs):
        r = pkg.call_tspws_main(gpu_lib, x, pkg.default_params(**kw))
        assert r["rc"] == 0
        np.testing.assert_array_equal(b["ls"][i], r["ls"])
        np.testing.assert_array_equal(b["tsPWS"][i], r["tsPWS"])
    o = O.tspws_main(O.Params(**kw), pairs[3])
    assert O.nerr(b["tsPWS"][3], o["tsPWS"]) <= TOL and O.nerr(b["ls"][3], o["ls"]) <= TOL
    # the parameter block comes back derived, like after tspws_main
    assert b["params"].J == o["params"].J and b["params"].V == o["params"].V


def test_batch_ragged_pairs_and_empty_batch(pkg, gpu_lib):
    """Pairs of different trace counts / lengths fall back to call-by-call; an empty batch is a no-op."""
    pairs = [O.synthetic_traces(12, 512, seed=1), O.synthetic_traces(30, 768, seed=2), O.synthetic_traces(7, 512, seed=3)]
    b = pkg.call_main_batch(gpu_lib, pairs, pkg.default_params(lrm=1))
    assert b["rc"] == 0
    for i, x in enumerate(pairs):
        o = O.tspws_main(O.Params(lrm=1), x)
        assert O.nerr(b["tsPWS"][i], o["tsPWS"]) <= TOL and O.nerr(b["ls"][i], o["ls"]) <= TOL
    assert pkg.call_main_batch(gpu_lib, [], pkg.default_params())["rc"] == 0
    assert gpu_lib.tspws_b200_main_batch(None, None, None, 3) == -1


def test_c5_station_pair_batch_named_size(pkg, gpu_lib):
    """configs[4] geometry: 500 traces x 8192 samples per pair, default Morlet (40 scales), nu = 2.
    12 pairs through the pipelined batch; two of them against the oracle, all of them against the
    linear stack by linearity (ls = synthesis(analysis(mean of the de-meaned traces)), one transform)."""
    pairs = _pairs(12, 500, 8192, 900)
    b = pkg.call_main_batch(gpu_lib, pairs, pkg.default_params(lrm=1))
    assert b["rc"] == 0
    assert b["params"].J * b["params"].V == 40
    fr = O.frame_from_params(O.derive_params(O.Params(lrm=1), 8192, 1.0), 8192)
    for i, x in enumerate(pairs):
        xo = x.copy()
        O.remove_mean_inplace(xo)
        lin = O.synthesis(fr, O.analysis(fr, xo.astype(np.float64).mean(axis=0)))
        assert O.nerr(b["ls"][i], lin) <= TOL
    for i in (0, 11):
        o = O.tspws_main(O.Params(lrm=1), pairs[i])
        assert O.nerr(b["tsPWS"][i], o["tsPWS"]) <= TOL


# ---------------------------------------------------------------------------------- resampling (C4)
def test_c4_resampling_named_size(pkg, gpu_lib):
    """configs[3]: 2000 x 32768, default Morlet (48 scales), M = 1000 sub-samples (p = 0.5), single-stage
    and two-stage (Kmax = 10, unbiased), jackknife n = 12, d = 2 (66 combinations)."""
    mtr, n, M = 2000, 32768, 1000
    x = O.synthetic_traces(mtr, n, seed=404)
    rng = np.random.default_rng(11)
    K = mtr // 2
    sel = np.zeros((M, mtr), np.int8)
    for m in range(M):
        sel[m, rng.choice(mtr, K, replace=False)] = 1
    xo = x.copy()
    O.remove_mean_inplace(xo)
    fr = O.frame_from_params(O.derive_params(O.Params(lrm=1), n, 1.0), n)
    assert fr.S == 48
    lin = (sel.astype(np.float32) @ xo) / np.float32(K)          # every realisation's linear stack

    gpu_lib.tspws_b200_set_selection(sel.ctypes.data_as(C.c_char_p), M, mtr)
    try:
        r1 = pkg.call_tspws_main(gpu_lib, x, pkg.default_params(lrm=1, subsmpl_N=M, subsmpl_p=0.5))
        r2 = pkg.call_tspws_main(gpu_lib, x, pkg.default_params(lrm=1, Kmax=10, unbiased=1, subsmpl_N=M, subsmpl_p=0.5))
    finally:
        gpu_lib.tspws_b200_set_selection(None, 0, 0)
    assert r1["rc"] == 0 and r2["rc"] == 0
    for r in (r1, r2):
        assert max(O.nerr(r["ls_subsmpl"][m], lin[m]) for m in range(M)) <= TOL
        assert np.isfinite(r["tsPWS_subsmpl"]).all()
    # two-stage realisations against the oracle (10 transforms each)
    rows = np.array([0, 499, 999])
    ls, ts, _ = O.resample_two_stage(fr, xo, sel[rows], 10, 2.0, True)
    for i, m in enumerate(rows):
        assert O.nerr(r2["tsPWS_subsmpl"][m], ts[i]) <= TOL and O.nerr(r2["ls_subsmpl"][m], ls[i]) <= TOL
    # single-stage: with every trace selected a realisation IS the main stack (weights (K,K) = (mtr,mtr))
    ones = np.ones((2, mtr), np.int8)
    gpu_lib.tspws_b200_set_selection(ones.ctypes.data_as(C.c_char_p), 2, mtr)
    try:
        r3 = pkg.call_tspws_main(gpu_lib, x, pkg.default_params(lrm=1, subsmpl_N=2, subsmpl_p=1.0))
    finally:
        gpu_lib.tspws_b200_set_selection(None, 0, 0)
    assert r3["rc"] == 0
    assert O.nerr(r3["tsPWS_subsmpl"][0], r3["tsPWS"]) <= 1e-5 and O.nerr(r3["tsPWS_subsmpl"][1], r1["tsPWS"]) <= 1e-5
    # jackknife from in->time
    times = CS.jackknife_times(mtr)
    r4 = pkg.call_tspws_main(gpu_lib, x, pkg.default_params(lrm=1, Kmax=10, unbiased=1, jackknife_n=12, jackknife_d=2), times=times)
    jsel = O.jackknife_plans(times, 12, 2)
    assert r4["rc"] == 0 and len(jsel) == 66
    np.testing.assert_array_equal(r4["mtr_subsmpl"], jsel.sum(axis=1))
    jr = np.array([0, 65])
    ls, ts, _ = O.resample_two_stage(fr, xo, jsel[jr], 10, 2.0, True)
    for i, m in enumerate(jr):
        assert O.nerr(r4["tsPWS_subsmpl"][m], ts[i]) <= TOL and O.nerr(r4["ls_subsmpl"][m], ls[i]) <= TOL
