"""Single-stage sub-sampling at the C4 trace shape (2000-sample-long filters, 32768 samples) with a reduced trace
and realisation count: the driver for `ncu -k regex:select_sum|resample_mma` captures.
    python tools/select_probe.py [mtr] [M]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import tspws_oracle as O   # noqa: E402
import conftest                        # noqa: E402

B = conftest.load_pkg()
lib = B.load_library()
mtr = int(sys.argv[1]) if len(sys.argv) > 1 else 256
M = int(sys.argv[2]) if len(sys.argv) > 2 else 128
n = 32768
x = O.synthetic_traces(mtr, n, seed=404)
for rep in range(2):
    t0 = time.time()
    r = B.call_tspws_main(lib, x, B.default_params(lrm=1, subsmpl_N=M, subsmpl_p=0.5))
    print(f"single-stage sub-sampling {mtr} x {n}, M = {M}: rc={r['rc']} {1e3 * (time.time() - t0):.1f} ms", flush=True)
