"""Worker of tests/test_gpu_parity.py::test_full_size_single_stage_w6_vs_oracle: ST / PS sums (float64) of a share
of the traces with the numpy oracle.  Test infrastructure; imports numpy and oracle/ only (never CUDA)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def partial_stacks(kw, n, traces):
    import numpy as np
    from oracle import tspws_oracle as O
    p = O.derive_params(O.Params(**kw), n, 1.0)
    fr = O.frame_from_params(p, n)
    st = ps = None
    for i in range(0, traces.shape[0], 64):          # bounded memory: 64 traces at a time
        a, b = O.stacks(fr, traces[i:i + 64].astype(np.float64))
        st = a if st is None else st + a
        ps = b if ps is None else ps + b
    return st, ps
