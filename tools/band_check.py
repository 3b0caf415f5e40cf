"""Banded tensor-core octaves (32 <= D < 256) vs oracle: per-octave coherence error, with the path on and off."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import tspws_oracle as O
import conftest, test_gpu_parity as T
B = conftest.load_pkg(); lib = B.load_library()
mtr, n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 16384
x = T._bench_traces(mtr, n)
for kw in (dict(w0=6.0, w0set=4), dict()):
    o = O.tspws_main(O.Params(**kw), x)
    fr = o["frame"]
    for tag, env in (("band+dense", None), ("dense only", "0")):
        if env is not None: os.environ["TSPWS_BAND_D"] = env
        a = T._plan_stack(B, lib, x, **kw)
        os.environ.pop("TSPWS_BAND_D", None)
        d = np.abs(a[2] - o["coherence"])
        print(kw, tag, "coh err", d.max(), "ts", O.nerr(a[1], o["tsPWS"]), "ls", O.nerr(a[0], o["ls"]))
        for D in sorted(set(fr.D.tolist())):
            idx = np.concatenate([np.arange(fr.coef_off[s], fr.coef_off[s + 1]) for s in range(fr.S) if fr.D[s] == D])
            print("     D", D, "max coh err %.2e" % d[idx].max())
