"""Where the time of the configs[3] resampling call goes: wall clock of tspws_main vs the in-stream kernel times."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import tspws_oracle as O
import conftest
B = conftest.load_pkg(); lib = B.load_library()
mtr, n, M = 2000, 32768, 1000
x = O.synthetic_traces(mtr, n, seed=404)
rng = np.random.default_rng(11)
sel = np.zeros((M, mtr), np.int8)
for m in range(M): sel[m, rng.choice(mtr, mtr // 2, replace=False)] = 1
lib.tspws_b200_set_selection(sel.ctypes.data_as(C.c_char_p), M, mtr)
for two_stage in (0, 1):
    kw = dict(lrm=1, subsmpl_N=M, subsmpl_p=0.5)
    if two_stage: kw.update(Kmax=10, unbiased=1)
    for rep in range(3):
        r = B.call_tspws_main(lib, x, B.default_params(**kw))
    print("two-stage" if two_stage else "single-stage", "wall ms (last of 3):", round(r["seconds"] * 1e3, 1), flush=True)
    lib.tspws_b200_profile(1); lib.tspws_b200_profile_reset()
    r = B.call_tspws_main(lib, x, B.default_params(**kw))
    prof = B.read_profile(lib); lib.tspws_b200_profile(0)
    tot = sum(v[1] for v in prof.values())
    print("   with kernel events: wall", round(r["seconds"] * 1e3, 1), "ms; kernels", round(tot, 1), "ms:",
          "  ".join(f"{k}={v[1]:.1f}ms(x{v[0]})" for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:10]), flush=True)
lib.tspws_b200_set_selection(None, 0, 0)
