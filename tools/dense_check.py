import os, sys, ctypes as C, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import tspws_oracle as O
import conftest, test_gpu_parity as T
B = conftest.load_pkg(); lib = B.load_library()
mtr, n = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 16384
x = T._bench_traces(mtr, n)
kw = dict(w0=6.0, w0set=4)
a = T._plan_stack(B, lib, x, **kw)
os.environ["TSPWS_DENSE_D"] = "0"
b = T._plan_stack(B, lib, x, **kw)
del os.environ["TSPWS_DENSE_D"]
print("dense vs simt: coh", O.nerr(a[2], b[2]), "ts", O.nerr(a[1], b[1]), "ls", O.nerr(a[0], b[0]))
d = np.abs(a[2] - b[2]); i = int(np.argmax(d)); print("worst coef", i, "of", len(d), a[2][i], b[2][i])
p = O.derive_params(O.Params(**kw), n, 1.0); fr = O.frame_from_params(p, n)
print("coef_off tail", fr.coef_off[-25:], "D", fr.D[-25:])
o = O.tspws_main(O.Params(**kw), x[:256]) if mtr <= 256 else None
