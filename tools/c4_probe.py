"""BASELINE configs[3] at full size on the GPU: 2000 traces x 32768 samples, default Morlet
(48 scales), M = 1000 sub-sampling realisations (single-stage and two-stage) + jackknife n=12, d=2.
Times every tspws_main call and checks a few realisations against the numpy oracle (explicit plans).

    python tools/c4_probe.py [rows-to-check]
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import cases as CS, tspws_oracle as O   # noqa: E402
import conftest                                      # noqa: E402

B = conftest.load_pkg()
lib = B.load_library()
ncheck = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mtr, n, M = 2000, 32768, 1000
x = O.synthetic_traces(mtr, n, seed=404)
rng = np.random.default_rng(11)
K = mtr // 2
sel = np.zeros((M, mtr), np.int8)
for m in range(M):
    sel[m, rng.choice(mtr, K, replace=False)] = 1


def run(tag, **kw):
    r = B.call_tspws_main(lib, x, B.default_params(**kw))
    dt = r["seconds"]
    print(f"{tag}: rc={r['rc']} {dt*1e3:.1f} ms in tspws_main (incl. H2D of {x.nbytes/1e6:.0f} MB)", flush=True)
    return r


# clean traces for the oracle (rm=1)
xo = x.copy()
O.remove_mean_inplace(xo)
p = O.derive_params(O.Params(lrm=1), n, 1.0)
fr = O.frame_from_params(p, n)
rows = np.linspace(0, M - 1, ncheck).astype(int)

lib.tspws_b200_set_selection(sel.ctypes.data_as(C.c_char_p), M, mtr)
run("warm-up single-stage main", lrm=1)
r1 = run("single-stage M=1000", lrm=1, subsmpl_N=M, subsmpl_p=0.5)
r1 = run("single-stage M=1000 (2nd)", lrm=1, subsmpl_N=M, subsmpl_p=0.5)
r2 = run("two-stage Kmax=10 M=1000", lrm=1, Kmax=10, unbiased=1, subsmpl_N=M, subsmpl_p=0.5)
r2 = run("two-stage Kmax=10 M=1000 (2nd)", lrm=1, Kmax=10, unbiased=1, subsmpl_N=M, subsmpl_p=0.5)
lib.tspws_b200_set_selection(None, 0, 0)
times = CS.jackknife_times(mtr)
t0 = time.time()
r3 = B.call_tspws_main(lib, x, B.default_params(lrm=1, Kmax=10, unbiased=1, jackknife_n=12, jackknife_d=2), times=times)
print(f"jackknife n=12 d=2 (C=66): rc={r3['rc']} {(time.time()-t0)*1e3:.1f} ms", flush=True)

t0 = time.time()
ls, ts = O.subsample_single_stage(fr, xo, sel[rows], 2.0, False)
print(f"oracle single-stage {ncheck} rows: {time.time()-t0:.1f} s")
for i, m in enumerate(rows):
    print(f"  row {m}: tsPWS err {O.nerr(r1['tsPWS_subsmpl'][m], ts[i]):.2e}  ls err {O.nerr(r1['ls_subsmpl'][m], ls[i]):.2e}")
ls, ts, _ = O.resample_two_stage(fr, xo, sel[rows], 10, 2.0, True)
for i, m in enumerate(rows):
    print(f"  two-stage row {m}: tsPWS err {O.nerr(r2['tsPWS_subsmpl'][m], ts[i]):.2e}  ls err {O.nerr(r2['ls_subsmpl'][m], ls[i]):.2e}")
jsel = O.jackknife_plans(times, 12, 2)
jr = np.array([0, len(jsel) - 1])
ls, ts, _ = O.resample_two_stage(fr, xo, jsel[jr], 10, 2.0, True)
for i, m in enumerate(jr):
    print(f"  jackknife row {m}: K={r3['mtr_subsmpl'][m]} (oracle {int(jsel[m].sum())}) tsPWS err {O.nerr(r3['tsPWS_subsmpl'][m], ts[i]):.2e}")
