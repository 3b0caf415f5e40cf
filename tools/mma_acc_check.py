import os, sys, ctypes as C, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest
B = conftest.load_pkg(); lib = B.load_library()
f = lib.tspws_cu_mma_selftest; f.restype = C.c_int
f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
rng = np.random.default_rng(1)
for T in (256, 1024, 4096, 16384):
    for nterm in (2, 3):
        M, ncoef = 128, 64
        cnt = np.ones((M, T), np.int8)
        # positive-mean data: accumulator grows, truncation bias would show as a systematic negative error
        P = (rng.uniform(0.5, 1.0, size=(T, ncoef, 2))).astype(np.float32)
        out = np.empty((M, 2 * ncoef), np.float32)
        assert f(cnt.ctypes.data, P.ctypes.data, M, T, ncoef, nterm, out.ctypes.data) == 0
        want = P.astype(np.float64).sum(axis=0).reshape(-1)
        rel = (out[0].astype(np.float64) - want) / want
        print(f"T={T:6d} nterm={nterm}: rel err mean {rel.mean():+.3e} max {np.abs(rel).max():.3e}")
