import os, sys, ctypes as C, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import tspws_oracle as O
import conftest, test_gpu_parity as T
B = conftest.load_pkg(); lib = B.load_library()
mtr, n = 256, 16384
x = T._bench_traces(mtr, n)
kw = dict(w0=6.0, w0set=4)
o = O.tspws_main(O.Params(**kw), x)
fr = o["frame"]
i0 = int(fr.coef_off[-21])      # first scale with D >= 256 (last 20 scales)
for tag, env in (("dense", None), ("simt", "0")):
    if env is not None: os.environ["TSPWS_DENSE_D"] = env
    a = T._plan_stack(B, lib, x, **kw)
    os.environ.pop("TSPWS_DENSE_D", None)
    d = np.abs(a[2] - o["coherence"])
    print(tag, "coh err all", d.max() / o["coherence"].max(), "coarse", d[i0:].max(), "fine", d[:i0].max(), "ts", O.nerr(a[1], o["tsPWS"]))
    for s in range(len(fr.coef_off) - 21, len(fr.coef_off) - 1):
        a0, a1 = int(fr.coef_off[s]), int(fr.coef_off[s + 1])
        print("   scale", s, "D", fr.D[s], "L", fr.L[s], "max err", d[a0:a1].max())
os.environ.pop("TSPWS_DENSE_D", None)
a = T._plan_stack(B, lib, x, **kw)
d = a[2] - o["coherence"]
a0 = int(fr.coef_off[39])
print("scale 39 errors x1e6:", np.round(d[a0:a0 + 64] * 1e6, 1))
print("coherence there:", np.round(o["coherence"][a0 + 20:a0 + 30], 4))
# per-trace check of the coefficient: oracle Y of every trace at coef 81560
Y = np.array([O.analysis(fr, x[t].astype(np.float64))[81560] for t in range(mtr)])
print("min |Y| over traces", np.abs(Y).min(), "median", np.median(np.abs(Y)), "5 smallest", np.sort(np.abs(Y))[:5])
