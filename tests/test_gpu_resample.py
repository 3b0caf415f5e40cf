"""Resampling (sub-sampling, jackknife) and convergence on the GPU, through tspws_main, against the
reference's outputs (tests/golden) with the same selection rows.
Tolerance: stacked traces <= 1e-4 (max|a-b| / max|b|), similarity curves <= 1e-5 absolute."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import cases as CS, tspws_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
TOL = 1e-4


@pytest.mark.parametrize("simt", [False, True], ids=["tensor_core_sums", "simt_fallbacks"])
@pytest.mark.parametrize("name", sorted(CS.RESAMPLE_CASES))
def test_resampling_matches_reference(pkg, gpu_lib, name, simt, monkeypatch):
    """Default: phase stacks, time-domain sums and group sums as tcgen05 contractions, per-trace phasors from the
    chunk-of-one transform kernels.  simt_fallbacks: the kernels they replaced (regroup_kernel, the per-trace transform),
    which odd trace lengths and very large Kmax still use."""
    if simt:
        monkeypatch.setenv("TSPWS_RS_REGROUP", "1")
        monkeypatch.setenv("TSPWS_RS_PERTRACE", "1")
    case = CS.RESAMPLE_CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    x = CS.make_input(case)
    sel = np.ascontiguousarray(g["sel"], dtype=np.int8)
    times = CS.jackknife_times(case["mtr"]) if case["plans"][0] == "jack" else None
    assert gpu_lib.tspws_b200_set_selection(sel.ctypes.data_as(C.c_char_p), sel.shape[0], sel.shape[1]) == 0
    try:
        r = pkg.call_tspws_main(gpu_lib, x, pkg.default_params(**case["params"]), dt=case["dt"], times=times)
    finally:
        gpu_lib.tspws_b200_set_selection(None, 0, 0)
    assert r["rc"] == 0
    assert O.nerr(r["tsPWS"], g["tsPWS"]) <= TOL and O.nerr(r["ls"], g["ls"]) <= TOL
    assert O.nerr(r["tsPWS_subsmpl"], g["tsPWS_subsmpl"]) <= TOL
    assert O.nerr(r["ls_subsmpl"], g["ls_subsmpl"]) <= TOL
    if case["plans"][0] == "jack":
        np.testing.assert_array_equal(r["mtr_subsmpl"], g["mtr_subsmpl"])


def test_jackknife_plans_from_times_without_override(pkg, gpu_lib):
    """Without an explicit selection the plans come from in->time like the reference's JackknifePlans."""
    name = "jackknife_two_stage_60x400"
    case = CS.RESAMPLE_CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    r = pkg.call_tspws_main(gpu_lib, CS.make_input(case), pkg.default_params(**case["params"]),
                            times=CS.jackknife_times(case["mtr"]))
    assert r["rc"] == 0
    np.testing.assert_array_equal(r["mtr_subsmpl"], g["mtr_subsmpl"])
    assert O.nerr(r["tsPWS_subsmpl"], g["tsPWS_subsmpl"]) <= TOL


def test_subsampling_rand_plans_follow_libc_sequence(pkg, gpu_lib):
    """Plans drawn inside tspws_main use the reference's rand() call sequence: seeding libc as in the
    fixture must reproduce the fixture's realisations."""
    name = "subsample_two_stage_24x600"
    case = CS.RESAMPLE_CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    C.CDLL(None).srand(case["plans"][1])
    r = pkg.call_tspws_main(gpu_lib, CS.make_input(case), pkg.default_params(**case["params"]))
    assert r["rc"] == 0
    assert O.nerr(r["tsPWS_subsmpl"], g["tsPWS_subsmpl"]) <= TOL
    assert O.nerr(r["ls_subsmpl"], g["ls_subsmpl"]) <= TOL


def test_single_stage_jackknife_is_a_stub_like_the_reference(pkg, gpu_lib):
    case = CS.RESAMPLE_CASES["jackknife_two_stage_60x400"]
    kw = dict(case["params"]); kw["Kmax"] = 0
    r = pkg.call_tspws_main(gpu_lib, CS.make_input(case), pkg.default_params(**kw), times=CS.jackknife_times(case["mtr"]))
    assert r["rc"] == 0
    assert not r["tsPWS_subsmpl"].any() and not r["ls_subsmpl"].any()     # reference: src/ts_pws1f_lib.c:711-716


@pytest.mark.parametrize("name", sorted(CS.CONVERGENCE_CASES))
def test_convergence_matches_reference(pkg, gpu_lib, name):
    case = CS.CONVERGENCE_CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    r = pkg.call_tspws_main(gpu_lib, CS.make_input(case), pkg.default_params(**case["params"]), dt=case["dt"])
    assert r["rc"] == 0
    assert O.nerr(r["tsPWS"], g["tsPWS"]) <= TOL
    for k in ("tsPWS_sim", "ls_sim"):
        assert np.max(np.abs(r[k] - g[k])) <= 1e-5
    for k in ("tsPWS_misfit", "ls_misfit"):
        assert O.nerr(r[k], g[k]) <= TOL
    if case["params"].get("AllSteps"):
        assert O.nerr(r["tsPWS_steps"], g["tsPWS_steps"]) <= TOL
        assert O.nerr(r["ls_steps"], g["ls_steps"]) <= TOL


def test_medium_two_stage_subsampling_vs_oracle(pkg, gpu_lib):
    """200 x 2048, M = 70 (> one batch), Kmax = 10: GPU vs numpy oracle with explicit plans."""
    rng = np.random.default_rng(3)
    x = O.synthetic_traces(200, 2048, seed=55)
    M, K = 70, 100
    sel = np.zeros((M, 200), np.int8)
    for m in range(M):
        sel[m, rng.choice(200, K, replace=False)] = 1
    kw = dict(lrm=1, Kmax=10, unbiased=1)
    gpu_lib.tspws_b200_set_selection(sel.ctypes.data_as(C.c_char_p), M, 200)
    try:
        r = pkg.call_tspws_main(gpu_lib, x, pkg.default_params(subsmpl_N=M, subsmpl_p=0.5, **kw))
    finally:
        gpu_lib.tspws_b200_set_selection(None, 0, 0)
    assert r["rc"] == 0
    o = O.tspws_main(O.Params(**kw), x)
    ls, ts, _ = O.resample_two_stage(o["frame"], o["sigall"], sel, 10, 2.0, True)
    assert O.nerr(r["tsPWS_subsmpl"], ts) <= TOL and O.nerr(r["ls_subsmpl"], ls) <= TOL


def test_single_stage_jackknife_extension(pkg, gpu_lib):
    """TSPWS_OPT_SS_JACKKNIFE: the single-stage delete-d jackknife the reference leaves as a stub
    (SURVEY.md 8f rank 3), against the oracle's single-stage resampling with per-row trace counts."""
    case = CS.RESAMPLE_CASES["jackknife_two_stage_60x400"]
    kw = dict(case["params"]); kw["Kmax"] = 0
    x = CS.make_input(case)
    times = CS.jackknife_times(case["mtr"])
    assert gpu_lib.tspws_b200_set_option(pkg.OPT_SS_JACKKNIFE, 1) == 0
    try:
        r = pkg.call_tspws_main(gpu_lib, x, pkg.default_params(**kw), times=times)
    finally:
        gpu_lib.tspws_b200_set_option(pkg.OPT_SS_JACKKNIFE, 0)
    assert r["rc"] == 0
    sel = O.jackknife_plans(times, kw["jackknife_n"], kw["jackknife_d"])
    np.testing.assert_array_equal(r["mtr_subsmpl"], sel.sum(axis=1))
    o = O.tspws_main(O.Params(**{k: v for k, v in kw.items() if k in ("lrm", "fold", "wu", "unbiased", "type")}), x)
    ls, ts = O.subsample_single_stage(o["frame"], o["sigall"], sel, o["params"].wu, bool(o["params"].unbiased))
    assert O.nerr(r["tsPWS_subsmpl"], ts) <= TOL and O.nerr(r["ls_subsmpl"], ls) <= TOL


def test_bootstrap_with_replacement_count_matrix(pkg, gpu_lib):
    """True bootstrap: the selection matrix holds COUNTS (traces drawn with replacement), the
    tensor-core contraction takes them as they are; single-stage stack against the oracle with the same counts."""
    mtr, n, M = 48, 1024, 12
    x = O.synthetic_traces(mtr, n, seed=77)
    rng = np.random.default_rng(5)
    cnt = np.zeros((M, mtr), np.int8)
    for m in range(M):
        np.add.at(cnt[m], rng.integers(0, mtr, size=mtr), 1)        # mtr draws with replacement
    assert cnt.max() > 1 and (cnt.sum(axis=1) == mtr).all()
    kw = dict(lrm=1, subsmpl_N=M, subsmpl_p=0.5)
    gpu_lib.tspws_b200_set_selection(cnt.ctypes.data_as(C.c_char_p), M, mtr)
    try:
        r = pkg.call_tspws_main(gpu_lib, x, pkg.default_params(**kw))
        assert r["rc"] == 0
        r2 = pkg.call_tspws_main(gpu_lib, x, pkg.default_params(Kmax=4, **kw))      # two-stage: counts are refused
        assert r2["rc"] != 0
    finally:
        gpu_lib.tspws_b200_set_selection(None, 0, 0)
    o = O.tspws_main(O.Params(lrm=1), x)
    ls, ts = O.subsample_single_stage(o["frame"], o["sigall"], cnt, 2.0, False)
    assert O.nerr(r["tsPWS_subsmpl"], ts) <= TOL and O.nerr(r["ls_subsmpl"], ls) <= TOL


def test_cli_resampling_and_convergence_outputs(gpu_lib):
    """tspws_b200_cli with convergence / AllSteps, jackknife + obin and sub-sampling flags: the files it writes
    (names and layouts of the reference CLI, src/ts_pws1f.c:354-418) equal tspws_main called directly."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cli_probe.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "CLI_PROBE_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
