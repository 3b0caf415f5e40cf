"""tcgen05 resampling contraction (ts-pws_b200/csrc/tspws_mma.cu) against numpy: counts [M x T] times the
unit-phasor matrix [T x ncoef] through the tile images and the UMMA kernel (internal hook tspws_cu_mma_selftest)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(lib, cnt, P, nterm):
    M, T = cnt.shape
    ncoef = P.shape[1]
    out = np.empty((M, 2 * ncoef), np.float32)
    f = lib.tspws_cu_mma_selftest
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    Pf = np.ascontiguousarray(np.stack([P.real, P.imag], axis=-1).astype(np.float32))
    rc = f(cnt.ctypes.data, Pf.ctypes.data, M, T, ncoef, nterm, out.ctypes.data)
    assert rc == 0, rc
    return out[:, 0::2] + 1j * out[:, 1::2], Pf


@pytest.mark.parametrize("M,T,ncoef,cmax,nterm", [(5, 70, 33, 1, 2), (130, 64, 64, 1, 2), (300, 200, 500, 1, 2),
                                                  (257, 333, 1000, 5, 2), (300, 200, 500, 3, 3), (1000, 520, 2100, 1, 2)])
def test_resample_mma_matches_numpy(gpu_lib, M, T, ncoef, cmax, nterm):
    rng = np.random.default_rng(M * 7 + T)
    cnt = rng.integers(0, cmax + 1, size=(M, T)).astype(np.int8)
    ph = rng.uniform(0, 2 * np.pi, size=(T, ncoef))
    P = np.exp(1j * ph)
    P[rng.random((T, ncoef)) < 0.01] = 0            # skipped (zero / non-finite) coefficients contribute nothing
    got, Pf = _run(gpu_lib, cnt, P, nterm)
    want = cnt.astype(np.float64) @ (Pf[..., 0].astype(np.float64) + 1j * Pf[..., 1])
    # bf16 terms: 2 terms keep 2^-17 of every phasor component, 3 terms 2^-25; fp32 accumulation on top
    per_term = 2.0 ** -17 if nterm == 2 else 2.0 ** -24
    bound = cnt.sum(axis=1).max() * per_term * 1.5 + 1e-6 * np.abs(want).max()
    assert np.abs(got - want).max() <= bound, (np.abs(got - want).max(), bound)
    # and statistically far better than the bound (random signs)
    assert np.abs(got - want).max() <= max(8 * np.sqrt(cnt.sum(axis=1).max()) * per_term, 2e-6 * np.abs(want).max() + 1e-6)
