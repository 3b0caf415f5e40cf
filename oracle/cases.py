"""Shared list of parity cases (inputs are synthetic and seeded, so they are regenerated, not stored).

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py (runs the compiled reference here and
stores its outputs under tests/golden/) and by tests/ (re-create the same inputs, compare)."""
from __future__ import annotations

import numpy as np

from . import tspws_oracle as O

# name -> dict(mtr, n, dt, seed, zero_trace (index or None), params (t_tsPWS fields))
CASES = {
    # the reference's own smoke recipe (src/Test_ts_pws.m:1-13): 30 x 1000, signal + noise, rm, wu=2, unbiased
    "recipe_30x1000": dict(mtr=30, n=1000, dt=1.0, seed=1, params=dict(lrm=1, wu=2.0, unbiased=1)),
    "morlet_default_24x2048": dict(mtr=24, n=2048, dt=1.0, seed=2, params=dict(lrm=1)),
    "morlet_w6_24x2048": dict(mtr=24, n=2048, dt=1.0, seed=3, params=dict(w0=6.0, w0set=4, lrm=1)),
    "mexhat_single_20x2048": dict(mtr=20, n=2048, dt=1.0, seed=4, params=dict(type=-3, lrm=1, fold=1)),
    "mexhat_two_stage_unbiased_40x2048": dict(mtr=40, n=2048, dt=1.0, seed=5,
                                              params=dict(type=-3, Kmax=10, unbiased=1, lrm=1, fold=1)),
    "morlet_two_stage_unbiased_64x4096": dict(mtr=64, n=4096, dt=1.0, seed=6,
                                              params=dict(Kmax=10, unbiased=1, lrm=1)),
    "morlet_two_stage_plain_48x2048": dict(mtr=48, n=2048, dt=1.0, seed=22, params=dict(Kmax=6, unbiased=1)),
    "morlet_two_stage_biased_37x1024": dict(mtr=37, n=1024, dt=1.0, seed=7, params=dict(Kmax=7, fold=1)),
    # example.sh configurations 2 and 3 in miniature: odd N, dt=4, fmin -> s0=7.89, D = 7/15/31
    "fmin_J3_oddN_25x1501": dict(mtr=25, n=1501, dt=4.0, seed=8, params=dict(fmin=0.004, J=3, lrm=1, fold=1)),
    "fmin_J3_oddN_two_stage_50x1501": dict(mtr=50, n=1501, dt=4.0, seed=9,
                                           params=dict(fmin=0.004, J=3, lrm=1, fold=1, Kmax=10, unbiased=1)),
    # edge cases
    "zero_trace_skipped_12x777": dict(mtr=12, n=777, dt=1.0, seed=10, zero_trace=5, params=dict()),
    "kmax1_unbiased_fallback_9x512": dict(mtr=9, n=512, dt=1.0, seed=11, params=dict(Kmax=1, unbiased=1)),
    "kmax_gt_mtr_single_9x512": dict(mtr=9, n=512, dt=1.0, seed=12, params=dict(Kmax=20, unbiased=1)),
    "nmax_7_of_16x512": dict(mtr=16, n=512, dt=1.0, seed=13, params=dict(Nmax=7, lrm=1)),
    "nu1_14x640": dict(mtr=14, n=640, dt=1.0, seed=14, params=dict(wu=1.0)),
    "nu3p5_14x640": dict(mtr=14, n=640, dt=1.0, seed=15, params=dict(wu=3.5, unbiased=1)),
    "uniform_sampling_10x300": dict(mtr=10, n=300, dt=1.0, seed=16, params=dict(uni=1, J=4)),
    "exact_morlet_10x900": dict(mtr=10, n=900, dt=1.0, seed=17, params=dict(type=-2, lrm=1)),
    "low_q_undecimated_8x400": dict(mtr=8, n=400, dt=1.0, seed=18, params=dict(w0=3.0, w0set=4)),
    "fold_refused_asymmetric_10x512": dict(mtr=10, n=512, dt=1.0, seed=19, beg=0.0, params=dict(fold=1)),
    "q_set_12x1000": dict(mtr=12, n=1000, dt=1.0, seed=20, params=dict(Q=4.0, w0set=1, lrm=1)),
    "single_trace_1x512": dict(mtr=1, n=512, dt=1.0, seed=21, params=dict()),
}


def make_input(case: dict) -> np.ndarray:
    x = O.synthetic_traces(case["mtr"], case["n"], seed=20171 + case["seed"])
    if case.get("zero_trace") is not None:
        x[case["zero_trace"]] = 0.0
    return x


# Resampling / convergence cases.  `plans`: how the reference's selection rows are reproduced:
#   ("rand", seed)  srand(seed) before the reference call; the same rows are re-drawn with the
#                   reference's SubsamplingPlan after re-seeding and stored in the fixture;
#   ("jack",)       JackknifePlans from synthetic daily start times.
RESAMPLE_CASES = {
    "subsample_two_stage_24x600": dict(mtr=24, n=600, dt=1.0, seed=31, plans=("rand", 99),
                                       params=dict(lrm=1, Kmax=4, unbiased=1, subsmpl_N=5, subsmpl_p=0.5)),
    "subsample_two_stage_fold_mexhat_30x512": dict(mtr=30, n=512, dt=1.0, seed=32, plans=("rand", 7),
                                       params=dict(type=-3, lrm=1, fold=1, Kmax=5, subsmpl_N=9, subsmpl_p=0.7)),
    "subsample_single_stage_16x512": dict(mtr=16, n=512, dt=1.0, seed=33, plans=("rand", 7),
                                       params=dict(lrm=1, subsmpl_N=4, subsmpl_p=0.5)),
    "subsample_single_stage_unbiased_20x700": dict(mtr=20, n=700, dt=1.0, seed=34, plans=("rand", 5),
                                       params=dict(lrm=1, unbiased=1, subsmpl_N=19, subsmpl_p=0.3)),
    "jackknife_two_stage_60x400": dict(mtr=60, n=400, dt=1.0, seed=35, plans=("jack",),
                                       params=dict(lrm=1, Kmax=4, unbiased=1, jackknife_n=6, jackknife_d=2)),
}
CONVERGENCE_CASES = {
    "convergence_single_12x400": dict(mtr=12, n=400, dt=1.0, seed=41, params=dict(lrm=1, unbiased=1, convergence=1, AllSteps=1)),
    "convergence_two_stage_14x400": dict(mtr=14, n=400, dt=1.0, seed=42, params=dict(lrm=1, Kmax=4, unbiased=1, convergence=1, AllSteps=1)),
    "convergence_mexhat_biased_9x300": dict(mtr=9, n=300, dt=1.0, seed=43, params=dict(type=-3, fold=1, Kmax=3, convergence=1)),
}


def jackknife_times(mtr: int) -> np.ndarray:
    """Synthetic daily start times from 2001-01-01, one trace every 5 days."""
    return (978307200 + 86400 * 5 * np.arange(mtr)).astype(np.int64)
