"""Pins the numpy oracle: against the golden vectors produced by the compiled reference
(tests/golden, made by oracle/gen_golden.py) and, where oracle/_ref is present, against the
reference library itself on fresh inputs."""
import hashlib
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import cases as CS, ref_lib as R, tspws_oracle as O

GOLD = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def oracle_run(case):
    x = CS.make_input(case)
    p = O.Params(**case["params"])
    return x, O.tspws_main(p, x, dt=case["dt"], beg=case.get("beg"))


@pytest.mark.parametrize("name", sorted(CS.CASES))
def test_oracle_matches_reference_golden(name):
    case = CS.CASES[name]
    g = load_golden(name)
    x, o = oracle_run(case)
    sha = np.frombuffer(hashlib.sha256(x.tobytes()).digest(), dtype=np.uint8)
    assert (sha == g["input_sha"]).all(), "synthetic input generator drifted from the fixtures"
    p = o["params"]
    np.testing.assert_array_equal(np.array([p.J, p.V, p.s0, p.b0, p.w0, p.fmin, p.fold], dtype=np.float64), g["derived"])
    assert o["K"] == int(g["K"])
    # float64 restatement vs float64 reference: agreement to rounding
    assert O.nerr(o["st"], g["st"]) < 1e-12
    assert O.nerr(o["ps"], g["ps"]) < 1e-11
    assert O.nerr(o["coherence"], np.abs(g["ps"]) / int(g["K"])) < 1e-11
    assert O.nerr(o["ls"], g["ls"]) < 2e-7         # float32 outputs: at most an ulp apart
    assert O.nerr(o["tsPWS"], g["tsPWS"]) < 2e-7


def test_direct_and_fft_analysis_agree():
    p = O.derive_params(O.Params(fmin=0.004, J=3), 301, 4.0)
    fr = O.frame_from_params(p, 301)
    x = O.synthetic_traces(1, 301, seed=5)[0]
    assert O.nerr(O.analysis(fr, x), O.analysis_direct(fr, x)) < 1e-13


def test_synthesis_is_adjoint_of_analysis():
    """<analysis(x), Z> = <x, synthesis-without-gain(Z)> -- checked through the gains."""
    p = O.derive_params(O.Params(), 512, 1.0)
    fr = O.frame_from_params(p, 512)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(512)
    Z = rng.standard_normal(fr.ncoef) + 1j * rng.standard_normal(fr.ncoef)
    gains = np.concatenate([np.full(fr.Ns[s], np.log(2.0) / (2 * fr.Cpsi * fr.V * fr.scale[s]) * fr.D[s])
                            for s in range(fr.S)])
    lhs = np.real(np.vdot(O.analysis(fr, x), Z * gains))     # Re sum conj(Y) Z g
    rhs = np.dot(x, O.synthesis(fr, Z))
    assert abs(lhs - rhs) < 1e-10 * max(abs(lhs), 1.0)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_live_reference_transforms():
    rf = R.RefFrame(-1, 5, 4, 1000, 2.0, 1.0, O.W0_DEFAULT, 0)
    fr = O.build_frame(-1, 5, 4, 1000, 2.0, 1.0, O.W0_DEFAULT, 0)
    x = O.synthetic_traces(1, 1000, seed=77)[0].astype(np.float64)
    Y = rf.analysis(x)
    assert O.nerr(O.analysis(fr, x), Y) < 1e-13
    assert O.nerr(O.synthesis(fr, Y), rf.synthesis(Y)) < 1e-13
    rf.close()


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_resampling_oracle_matches_live_reference(pkg):
    """Two-stage sub-sampling with the reference's own rand() plans (OMP_NUM_THREADS=1 semantics:
    plans are drawn serially here and the same plans are given to the oracle)."""
    import ctypes as C
    ref = pkg.declare_reference(R.load())
    libc = C.CDLL(None)
    x = O.synthetic_traces(24, 600, seed=31)
    M, prob, Kmax = 5, 0.5, 4
    C.CDLL("libgomp.so.1").omp_set_num_threads(1)    # rand() inside `omp for` (src/ts_pws1f_lib.c:640)
    p = pkg.default_params(lrm=1, Kmax=Kmax, unbiased=1, subsmpl_N=M, subsmpl_p=prob)
    libc.srand(99)
    r = pkg.call_tspws_main(ref, x, p)
    libc.srand(99)
    K = int(np.ceil(24 * prob))
    sel = R.subsampling_plans(M, 24, K)
    po = O.Params(lrm=1, Kmax=Kmax, unbiased=1)
    o = O.tspws_main(po, x)
    ls, ts, Ks = O.resample_two_stage(o["frame"], o["sigall"], sel, Kmax, 2.0, True)
    assert O.nerr(ts, r["tsPWS_subsmpl"]) < 1e-6
    assert O.nerr(ls, r["ls_subsmpl"]) < 1e-6


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_single_stage_resampling_oracle_matches_live_reference(pkg):
    import ctypes as C
    ref = pkg.declare_reference(R.load())
    libc = C.CDLL(None)
    x = O.synthetic_traces(16, 512, seed=32)
    M, prob = 4, 0.5
    p = pkg.default_params(lrm=1, subsmpl_N=M, subsmpl_p=prob)
    libc.srand(7)
    r = pkg.call_tspws_main(ref, x, p)
    libc.srand(7)
    sel = R.subsampling_plans(M, 16, int(np.ceil(16 * prob)))
    o = O.tspws_main(O.Params(lrm=1), x)
    ls, ts = O.subsample_single_stage(o["frame"], o["sigall"], sel, 2.0, False)
    assert O.nerr(ts, r["tsPWS_subsmpl"]) < 1e-6
    assert O.nerr(ls, r["ls_subsmpl"]) < 1e-6
