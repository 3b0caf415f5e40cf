"""Pins the resampling / convergence oracle against what the compiled reference produced
(tests/golden, oracle/gen_golden.py): same selection rows, same traces."""
import hashlib
import math
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import cases as CS, tspws_oracle as O

GOLD = os.path.join(ROOT, "tests", "golden")


def _check_input(x, g):
    sha = np.frombuffer(hashlib.sha256(x.tobytes()).digest(), dtype=np.uint8)
    assert (sha == g["input_sha"]).all()


@pytest.mark.parametrize("name", sorted(CS.RESAMPLE_CASES))
def test_resampling_oracle_matches_reference(name):
    case = CS.RESAMPLE_CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    x = CS.make_input(case)
    _check_input(x, g)
    kw = {k: v for k, v in case["params"].items() if k in O.Params.__dataclass_fields__}
    o = O.tspws_main(O.Params(**kw), x, dt=case["dt"])
    sel = g["sel"]
    p = o["params"]
    if p.Kmax and p.Kmax <= case["mtr"]:
        ls, ts, Ks = O.resample_two_stage(o["frame"], o["sigall"], sel, p.Kmax, p.wu, bool(p.unbiased))
        if "jackknife_n" in case["params"]:
            np.testing.assert_array_equal(Ks, g["mtr_subsmpl"])
            np.testing.assert_array_equal(sel, O.jackknife_plans(CS.jackknife_times(case["mtr"]),
                                                                 case["params"]["jackknife_n"], case["params"]["jackknife_d"]))
    else:
        ls, ts = O.subsample_single_stage(o["frame"], o["sigall"], sel, p.wu, bool(p.unbiased))
    assert O.nerr(ts, g["tsPWS_subsmpl"]) < 1e-6
    assert O.nerr(ls, g["ls_subsmpl"]) < 1e-6


@pytest.mark.parametrize("name", sorted(CS.CONVERGENCE_CASES))
def test_convergence_oracle_matches_reference(name):
    case = CS.CONVERGENCE_CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    x = CS.make_input(case)
    _check_input(x, g)
    kw = {k: v for k, v in case["params"].items() if k in O.Params.__dataclass_fields__}
    c = O.convergence(O.Params(**kw), x, dt=case["dt"])
    for k in ("tsPWS_sim", "ls_sim"):
        assert np.max(np.abs(c[k] - g[k])) < 1e-6
    for k in ("tsPWS_misfit", "ls_misfit"):
        assert O.nerr(c[k], g[k]) < 1e-5
    if "tsPWS_steps" in g:
        assert O.nerr(c["tsPWS_steps"], g["tsPWS_steps"]) < 1e-6
        assert O.nerr(c["ls_steps"], g["ls_steps"]) < 1e-6
