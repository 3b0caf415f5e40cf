import os, sys, ctypes as C, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest
B = conftest.load_pkg(); lib = B.load_library()
f = lib.tspws_cu_mma_selftest; f.restype = C.c_int
f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
rng = np.random.default_rng(1)
for T in (64, 256, 1024, 4096, 16384):
    for nterm in (2, 3):
        M, ncoef = 128, 64
        cnt = np.ones((M, T), np.int8)
        P = rng.standard_normal(size=(T, ncoef, 2)).astype(np.float32)      # random signs: sum ~ sqrt(T)
        out = np.empty((M, 2 * ncoef), np.float32)
        assert f(cnt.ctypes.data, P.ctypes.data, M, T, ncoef, nterm, out.ctypes.data) == 0
        want = P.astype(np.float64).sum(axis=0).reshape(-1)
        err = out[0].astype(np.float64) - want
        print(f"T={T:6d} nterm={nterm}: rms err / rms sum {np.sqrt((err**2).mean()) / np.sqrt((want**2).mean()):.3e}   (rms err {np.sqrt((err**2).mean()):.3e}, sqrt(T) {np.sqrt(T):.1f})")
