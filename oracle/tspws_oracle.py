"""CPU restatement (numpy, float64) of the reference ts-PWS stacking path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; the product
(ts-pws_b200/csrc -> libtspws_b200.so) never does and has no CPU fallback.

Parity status: PINNED -- the reference ships no golden vectors, so
this restatement is pinned against the reference ITSELF compiled here from
/root/reference/src into oracle/_ref (oracle/Makefile): tests/test_oracle_golden.py and
tests/test_oracle_resample.py when oracle/_ref is present, and tests/golden/*.npz (made by oracle/gen_golden.py from
the compiled reference) everywhere else.

Everything is written from the mathematics; citations give the
reference lines each function restates.  The analysis / synthesis use the exact FFT
identity of the reference's circular time-domain correlations.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

PI = 3.14159265358979328            # the reference's own constant (src/ts_pws1f_lib.h:15-17)
W0_DEFAULT = PI * math.sqrt(2.0 / math.log(2.0))
NSIGMAS = 5                         # src/FWTa/wavelet_v7.h:20


# --------------------------------------------------------------------------------------
# Parameters
# --------------------------------------------------------------------------------------
@dataclass
class Params:
    """The fields of t_tsPWS the library acts on (src/ts_pws1f_lib.h:19-57), CLI defaults
    (src/ts_pws1f.c:140-142)."""
    type: int = -1
    uni: int = 0
    J: int = 0
    V: int = 4
    s0: float = 2.0
    b0: float = 1.0
    w0: float = W0_DEFAULT
    wu: float = 2.0
    fmin: float = 0.0
    Q: float = 0.0
    cycle: float = 2.0
    w0set: int = 0
    lrm: int = 0
    lVfix: int = 0
    ls0fix: int = 0
    lb0fix: int = 0
    fold: int = 0
    unbiased: int = 0
    Nmax: int = 0
    Kmax: int = 0


def derive_params(p: Params, nsamp: int, dt: float) -> Params:
    """Restates src/ts_pws1f_lib.c:91-124 (mutates and returns p). dt is a C float there."""
    dt = float(np.float32(dt))
    if p.fmin != 0 and p.fmin < 1.0 / (dt * nsamp):
        p.fmin = 0.0
    if p.w0set == 1:
        p.w0 = 2.0 * math.sqrt(math.log(2.0)) * p.Q
    elif p.w0set == 2:
        p.w0 = PI / math.sqrt(math.log(2.0)) * p.cycle
    if p.type in (-1, -2):
        r = p.w0 / (PI * math.sqrt(2.0 / math.log(2.0)))
        if not p.lVfix:
            p.V = int(math.ceil(4.0 * r))
        if not p.lb0fix:
            # C round() = half away from zero; the value passes through (unsigned int)
            p.b0 = float(int(2.0 ** _c_round(math.log2(r))))
        if not p.ls0fix:
            p.s0 = 2.0
    if p.type == -3:
        p.w0 = math.sqrt(2.0)
        if not p.lVfix:
            p.V = 2
        if not p.lb0fix:
            p.b0 = 0.5
        if not p.ls0fix:
            p.s0 = 1.0
    if p.fmin:
        a = p.w0 / (2.0 * PI * dt * p.fmin)
        if p.J:
            a /= 2.0 ** (p.J - 1.0 / p.V)
            while a < p.s0 * 0.9:
                a *= 2.0
                p.J -= 1
            p.s0 = a
        else:
            p.J = int(math.floor(math.log2(a / p.s0) + 1.0 / p.V))
    elif not p.J:
        a = nsamp * p.w0 / (2.0 * PI * 4.0 * p.s0)
        p.J = int(math.floor(math.log2(a) + 1.0 / p.V))
    return p


def _c_round(x: float) -> float:
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


# --------------------------------------------------------------------------------------
# Frame (filter bank)
# --------------------------------------------------------------------------------------
@dataclass
class Frame:
    type: int
    N: int
    V: int
    S: int
    Cpsi: float
    scale: np.ndarray
    L: np.ndarray
    center: np.ndarray
    D: np.ndarray
    Ns: np.ndarray
    w: list = field(default_factory=list)   # per-scale complex128 filters

    @property
    def ncoef(self) -> int:
        return int(self.Ns.sum())

    @property
    def coef_off(self) -> np.ndarray:
        return np.concatenate([[0], np.cumsum(self.Ns)])


def _erfi_series(z: np.ndarray) -> np.ndarray:
    """Maclaurin series, 500 terms, as src/FWTa/wavelet_def_v7.c:104-117."""
    z = np.asarray(z, dtype=np.float64)
    zz = z * z
    term = z.copy()
    out = term.copy()
    for n in range(1, 500):
        term = term * (zz / n)
        out = out + term / (2 * n + 1)
    return out * (2.0 / math.sqrt(PI))


def morlet_cpsi(w0: float) -> float:
    """src/FWTa/wavelet_def_v7.c:133-144 -- the loop variable accumulates 0.01 in floating
    point, so the same running sum is used here."""
    total = 0.0
    w = 0.01
    while w < 100:
        a = (w - w0)
        total += math.exp(-a * a) / w
        w += 0.01
    return total * 0.01 * math.sqrt(PI) / 2.0


def build_frame(type_: int, J: int, V: int, N: int, s0: float, b0: float, w0: float,
                uni: int = 0) -> Frame:
    """Restates CreateWaveletFamily and helpers (src/FWTa/wavelet_def_v7.c:204-339)."""
    S = J * V
    # scales by repeated multiplication (:299-309)
    scale = np.empty(S)
    a = s0
    step = 2.0 ** (1.0 / V)
    for s in range(S):
        scale[s] = a
        a *= step
    # lengths (:312-322)
    L = np.array([min(N, 2 * int(math.ceil(NSIGMAS * sc)) + 1) for sc in scale], dtype=np.int64)
    center = L // 2
    # decimation per octave (:325-339); b0 := 1 when uniform (:253)
    D = np.ones(S, dtype=np.int64)
    if not uni:
        a = s0 * b0
        for j in range(J):
            D[j * V:(j + 1) * V] = 1 if a < 1 else int(a)
            a *= 2.0
    Ns = (N + D - 1) // D
    w = []
    for s in range(S):
        u = (np.arange(L[s]) - center[s]) / scale[s]
        if type_ == -1:      # Morlet (:38-52)
            k = 1.0 / math.sqrt(math.sqrt(PI) * scale[s])
            w.append(k * np.exp(1j * w0 * u) * np.exp(-0.5 * u * u))
        elif type_ == -2:    # exact Morlet (:71-87)
            k = 1.0 / math.sqrt(math.sqrt(PI) * scale[s])
            w.append(k * np.exp(-0.5 * u * u) * (np.exp(1j * w0 * u) - math.exp(-w0 * w0 / 2)))
        elif type_ == -3:    # complex Mexican hat (:119-131)
            k = 2.0 / math.sqrt(3.0 * math.sqrt(PI) * scale[s])
            w.append(k * ((u * u - 1) * np.exp(-u * u / 2) * (1 + 1j * _erfi_series(u / math.sqrt(2.0)))
                          - 1j * math.sqrt(2.0 / PI) * u))
        else:
            raise ValueError("type must be -1, -2 or -3")
    Cpsi = morlet_cpsi(w0) if type_ in (-1, -2) else (4.0 / 3.0) * math.sqrt(PI)
    return Frame(type_, N, V, S, Cpsi, scale, L, center, D, Ns, w)


def frame_from_params(p: Params, N: int) -> Frame:
    return build_frame(p.type, p.J, p.V, N, p.s0, p.b0, p.w0, p.uni)


# --------------------------------------------------------------------------------------
# Analysis / synthesis
# --------------------------------------------------------------------------------------
def _filter_spectra(fr: Frame):
    """FFT of the circularly placed filters h_s[(l - c_s) mod N] = w_s[l]."""
    if getattr(fr, "_H", None) is None:
        H = []
        for s in range(fr.S):
            h = np.zeros(fr.N, dtype=np.complex128)
            idx = (np.arange(fr.L[s]) - fr.center[s]) % fr.N
            np.add.at(h, idx, fr.w[s])
            H.append(np.fft.fft(h))
        fr._H = H
    return fr._H


def analysis(fr: Frame, x: np.ndarray) -> np.ndarray:
    """Y_s[m] = sum_l x[(m D_s - c_s + l) mod N] conj(w_s[l])  (src/FWTa/wavelet_v7.c:43-64,
    src/FWTa/cdotx.c:35-72).  x: [..., N] real.  Returns ragged [..., ncoef] complex128."""
    x = np.asarray(x, dtype=np.float64)
    X = np.fft.fft(x, axis=-1)
    H = _filter_spectra(fr)
    out = np.empty(x.shape[:-1] + (fr.ncoef,), dtype=np.complex128)
    off = fr.coef_off
    for s in range(fr.S):
        # sum_l x[n+l-c] conj(w[l]) = IFFT( X . conj(FFT(h)) ) with h[(l-c)%N] = w[l]
        y = np.fft.ifft(X * np.conj(H[s]), axis=-1)
        out[..., off[s]:off[s + 1]] = y[..., ::fr.D[s]]
    return out


def analysis_direct(fr: Frame, x: np.ndarray) -> np.ndarray:
    """Same as analysis() by the literal time-domain sum (small cases / cross-check)."""
    x = np.asarray(x, dtype=np.float64)
    N = fr.N
    out = np.empty(fr.ncoef, dtype=np.complex128)
    off = fr.coef_off
    for s in range(fr.S):
        l = np.arange(fr.L[s])
        for m in range(fr.Ns[s]):
            idx = (m * fr.D[s] - fr.center[s] + l) % N
            out[off[s] + m] = np.sum(x[idx] * np.conj(fr.w[s]))
    return out


def synthesis(fr: Frame, Z: np.ndarray) -> np.ndarray:
    """x^[n] = sum_s a_s D_s sum_m Re( Z_s[m] w_s[n - m D_s + c_s] ), a_s = ln2/(2 Cpsi V scale_s)
    (src/FWTa/wavelet_v7.c:124-150, src/FWTa/cdotx.c:305-340 and :176-211).
    Z: ragged [..., ncoef] complex.  Returns [..., N] float64."""
    Z = np.asarray(Z, dtype=np.complex128)
    N = fr.N
    H = _filter_spectra(fr)
    off = fr.coef_off
    acc = np.zeros(Z.shape[:-1] + (N,), dtype=np.complex128)
    for s in range(fr.S):
        up = np.zeros(Z.shape[:-1] + (N,), dtype=np.complex128)
        up[..., (np.arange(fr.Ns[s]) * fr.D[s])] = Z[..., off[s]:off[s + 1]]
        a = math.log(2.0) / (2.0 * fr.Cpsi * fr.V * fr.scale[s]) * fr.D[s]
        # sum_p up[p] h[(n-p) mod N]  (circular convolution)
        acc += a * np.fft.ifft(np.fft.fft(up, axis=-1) * H[s], axis=-1)
    return acc.real


# --------------------------------------------------------------------------------------
# Stacks
# --------------------------------------------------------------------------------------
def fold_inplace(x: np.ndarray) -> None:
    """src/ts_pws1f_lib.c:76-86, float32 arithmetic, in place.  x: [mtr, max] float32."""
    n = x.shape[1]
    h = n // 2
    a = (x[:, :h] + x[:, ::-1][:, :h]).astype(np.float32) * np.float32(0.5)
    x[:, :h] = a
    x[:, n - h:] = a[:, ::-1]


def fold_allowed(beg: float, dt: float, nmax: int) -> bool:
    """src/ts_pws1f_lib.c:72 (float arithmetic promoted to double by the 0.5 literal)."""
    beg, dt = np.float32(beg), np.float32(dt)
    lhs = np.float32(np.float32(2) * beg) + np.float32(np.float32(nmax - 1) * dt)
    return not (float(lhs) > 0.5 * float(dt))


def remove_mean_inplace(x: np.ndarray) -> None:
    """src/ts_pws1f_lib.c:160-168: double mean, cast to float, float subtraction."""
    m = (x.astype(np.float64).sum(axis=1) / x.shape[1]).astype(np.float32)
    x -= m[:, None]


def group_index(mtr: int, K: int) -> np.ndarray:
    """floor(itr*Kmax/mtr) (src/ts_pws1f_lib.c:876), computed in unsigned int then double."""
    itr = np.arange(mtr, dtype=np.int64)
    return np.floor(((itr * K) % (1 << 32)).astype(np.float64) / float(mtr)).astype(np.int64)


def partial_linear_stacks(x: np.ndarray, K: int) -> np.ndarray:
    """src/ts_pws1f_lib.c:866-881 -> [K, max] float64 group sums."""
    mtr = x.shape[0]
    g = group_index(mtr, K)
    out = np.zeros((K, x.shape[1]))
    np.add.at(out, g, x.astype(np.float64))
    return out


def phasor(Y: np.ndarray) -> np.ndarray:
    """c/|c| with the reference's guard (src/ts_pws1f_lib.c:491-492): non-finite results
    (zero coefficient) are skipped."""
    a = np.abs(Y)
    with np.errstate(divide="ignore", invalid="ignore"):
        ph = Y / a
    bad = ~np.isfinite(ph.real) | ~np.isfinite(ph.imag) | ~((ph * np.conj(ph)).real <= 1.001)
    ph[bad] = 0
    return ph


def stacks(fr: Frame, traces: np.ndarray, batch: int = 64):
    """ST = sum_t Y_t, PS = sum_t Y_t/|Y_t| (src/ts_pws1f_lib.c:466-499 / :885-906)."""
    st = np.zeros(fr.ncoef, dtype=np.complex128)
    ps = np.zeros(fr.ncoef, dtype=np.complex128)
    for i in range(0, traces.shape[0], batch):
        Y = analysis(fr, traces[i:i + batch])
        st += Y.sum(axis=0)
        ps += phasor(Y).sum(axis=0)
    return st, ps


def weight(st: np.ndarray, ps: np.ndarray, K: int, mtr: int, wu: float, unbiased: bool) -> np.ndarray:
    """tspws_biased (src/ts_pws1f_lib.c:909-943) / tspws_unbiased (:965-984); the caller's
    rule `wu == 2 && unbiased` (:226) is applied here."""
    a2 = ps.real ** 2 + ps.imag ** 2
    if wu == 2 and unbiased and K > 1:
        c = a2 * (1.0 / K) ** 2
        return st * ((K * c - 1.0) * (1.0 / (K - 1))) * (1.0 / mtr)
    if wu == 2:
        return st * (a2 * (1.0 / (float(K) * float(K) * float(mtr))))
    if wu == 1:
        return st * np.sqrt(a2) * (1.0 / (float(K) * float(mtr)))
    return st * (np.sqrt(a2) / K) ** wu / mtr


def tspws_main(p: Params, sigall: np.ndarray, dt: float = 1.0, beg: Optional[float] = None) -> dict:
    """Restates tspws_main for the main stack (src/ts_pws1f_lib.c:48-243).  Returns ls, tsPWS
    (float32), the coherence map |PS|/K, ST, PS, the frame and the mutated params/traces."""
    x = np.array(sigall, dtype=np.float32, copy=True)
    nmax = x.shape[1]
    if beg is None:
        beg = -(nmax - 1) * dt / 2.0
    mtr = p.Nmax if p.Nmax else x.shape[0]
    x = x[:mtr]
    if p.fold:
        if fold_allowed(beg, dt, nmax):
            fold_inplace(x)
        else:
            p.fold = 0
    derive_params(p, nmax, dt)
    fr = frame_from_params(p, nmax)
    if p.lrm:
        remove_mean_inplace(x)
    if not p.Kmax or p.Kmax > mtr:
        K = mtr
        st, ps = stacks(fr, x)
    else:
        K = p.Kmax
        st, ps = stacks(fr, partial_linear_stacks(x, K))
    out = weight(st, ps, K, mtr, p.wu, bool(p.unbiased))
    ts = synthesis(fr, out).astype(np.float32)
    ls = (synthesis(fr, st).astype(np.float32) / np.float32(mtr)).astype(np.float32)
    return dict(ls=ls, tsPWS=ts, coherence=np.abs(ps) / K, st=st, ps=ps, K=K, mtr=mtr,
                frame=fr, params=p, sigall=x)


# --------------------------------------------------------------------------------------
# Resampling
# --------------------------------------------------------------------------------------
def subsample_single_stage(fr: Frame, x: np.ndarray, sel: np.ndarray, wu: float, unbiased: bool):
    """tspws_subsmpl_float (src/ts_pws1f_lib.c:501-610) for explicit sel[M, mtr] (0/1):
    ST_m = Sel.Y, PS_m = Sel.phasor(Y); weights with (K, K); ls_m in the TIME domain with
    float32 accumulators (:537-542, :579-583).  x: [mtr, max] float32 (already folded / rm)."""
    M, mtr = sel.shape
    cnt = np.where(sel > 0, sel, 0).astype(np.int64)    # 0/1 as in the reference; entries > 1 = bootstrap counts (extension)
    Ks = cnt.sum(axis=1)                   # the reference has one K for all rows; the per-row form also
    st = np.zeros((M, fr.ncoef), dtype=np.complex128)   # serves the single-stage jackknife extension
    ps = np.zeros((M, fr.ncoef), dtype=np.complex128)
    ls = np.zeros((M, x.shape[1]), dtype=np.float32)
    for t in range(mtr):
        Y = analysis(fr, x[t])
        ph = phasor(Y)
        for c in range(1, int(cnt[:, t].max()) + 1):    # a trace drawn c times is added c times
            use = cnt[:, t] >= c
            st[use] += Y
            ps[use] += ph
            ls[use] = (ls[use] + x[t].astype(np.float64)).astype(np.float32)   # float += double
    ts = np.empty_like(ls)
    for m in range(M):
        K = int(Ks[m])
        ls[m] = (ls[m] * np.float32(1.0 / K)).astype(np.float32)
        ts[m] = synthesis(fr, weight(st[m], ps[m], K, K, wu, unbiased)).astype(np.float32)
    return ls, ts


def regroup(x: np.ndarray, sel_row: np.ndarray, Kmax: int) -> np.ndarray:
    """Selected traces, by rank r within the selection, go to group floor(r*Kmax/K)
    (src/ts_pws1f_lib.c:644-654, :759-772)."""
    idx = np.nonzero(sel_row == 1)[0]
    K = len(idx)
    g = np.floor((np.arange(K, dtype=np.int64) * Kmax).astype(np.float64) / float(K)).astype(np.int64)
    out = np.zeros((Kmax, x.shape[1]))
    np.add.at(out, g, x[idx].astype(np.float64))
    return out


def resample_two_stage(fr: Frame, x: np.ndarray, sel: np.ndarray, Kmax: int, wu: float, unbiased: bool):
    """TwoStage_subsmpl_float / TwoStage_jackknife_float (src/ts_pws1f_lib.c:612-709, :719-831)
    for explicit sel[M, mtr].  Returns ls[M,max], tsPWS[M,max] float32 and K[M]."""
    M = sel.shape[0]
    ls = np.empty((M, x.shape[1]), dtype=np.float32)
    ts = np.empty_like(ls)
    Ks = np.empty(M, dtype=np.int64)
    for m in range(M):
        G = regroup(x, sel[m], Kmax)
        K = int((sel[m] == 1).sum())
        st, ps = stacks(fr, G)
        ts[m] = synthesis(fr, weight(st, ps, Kmax, K, wu, unbiased)).astype(np.float32)
        ls[m] = (G.sum(axis=0) * (1.0 / K)).astype(np.float32)
        Ks[m] = K
    return ls, ts, Ks


def jackknife_plans(times: np.ndarray, n: int, d: int) -> np.ndarray:
    """JackknifePlans (src/ts_pws1f_lib.c:385-430): bins floor(yday*n/365) from gmtime, all
    C(n,d) delete-d combinations in lexicographic order -> sel[C, mtr] (1 = kept)."""
    import itertools
    import time as _time
    yday = np.array([_time.gmtime(int(t)).tm_yday - 1 for t in times])    # tm_yday is 0-based in C
    bins = np.floor((yday * n).astype(np.float64) / 365.0).astype(np.int64)
    combos = list(itertools.combinations(range(n), d))
    sel = np.ones((len(combos), len(times)), dtype=np.int8)
    for c, comb in enumerate(combos):
        for b in comb:
            sel[c, bins == b] = 0
    return sel


# --------------------------------------------------------------------------------------
# Convergence helpers
# --------------------------------------------------------------------------------------
def similarity(x1: np.ndarray, x2: np.ndarray) -> float:
    """GNCC, src/ts_pws1f_lib.c:433-449."""
    x1 = np.asarray(x1, np.float64)
    x2 = np.asarray(x2, np.float64)
    return float(np.dot(x1, x2) / math.sqrt(np.dot(x1, x1)) / math.sqrt(np.dot(x2, x2)))


def misfit(x1: np.ndarray, x2: np.ndarray) -> float:
    """src/ts_pws1f_lib.c:452-462."""
    d = np.asarray(x1, np.float64) - np.asarray(x2, np.float64)
    return float(np.dot(d, d))


def convergence(p: Params, sigall: np.ndarray, dt: float = 1.0, reference=None) -> dict:
    """Convergence curves (src/ts_pws1f_lib.c:247-314) for already-derived p: after every added trace
    the ts-PWS of the first Tr traces (plain stack while Tr <= Kmax or Kmax == 0, regrouped two-stage
    stack otherwise) and the running time-domain mean, each compared (GNCC, squared misfit) with the
    final stack or `reference`."""
    main = tspws_main(p, sigall, dt=dt)
    x, fr, mtr = main["sigall"], main["frame"], main["mtr"]
    ref_ts = reference if reference is not None else main["tsPWS"]
    ref_ls = reference if reference is not None else main["ls"]
    ts_sim, ts_mis, ls_sim, ls_mis = (np.zeros(mtr) for _ in range(4))
    ts_steps = np.zeros((mtr, x.shape[1]), np.float32)
    ls_steps = np.zeros((mtr, x.shape[1]), np.float32)
    st = np.zeros(fr.ncoef, complex)
    ps = np.zeros(fr.ncoef, complex)
    for itr in range(mtr):
        Tr = itr + 1
        if not p.Kmax or p.Kmax >= Tr:
            K = Tr
            Y = analysis(fr, x[itr])
            st = st + Y
            ps = ps + phasor(Y)
        else:
            K = p.Kmax
            st, ps = stacks(fr, partial_linear_stacks(x[:Tr], K))
        d = synthesis(fr, weight(st, ps, K, Tr, p.wu, bool(p.unbiased)))
        ts_sim[itr], ts_mis[itr] = similarity(d, ref_ts), misfit(d, ref_ts)
        ts_steps[itr] = d.astype(np.float32)
    run = np.zeros(x.shape[1])
    for itr in range(mtr):
        run = run + x[itr].astype(np.float64)
        fa1 = float(np.float32(1.0 / (itr + 1)))
        run = run * fa1
        ls_sim[itr], ls_mis[itr] = similarity(run, ref_ls), misfit(run, ref_ls)
        ls_steps[itr] = run.astype(np.float32)
        run = run * (itr + 1)
    return dict(main=main, tsPWS_sim=ts_sim, tsPWS_misfit=ts_mis, ls_sim=ls_sim, ls_misfit=ls_mis,
                tsPWS_steps=ts_steps, ls_steps=ls_steps)


# --------------------------------------------------------------------------------------
# Synthetic traces -- counter-based so any rank can make any trace
# --------------------------------------------------------------------------------------
def synthetic_signal(nmax: int) -> np.ndarray:
    """Windowed exponential chirp in the style of the reference's src/FWTa/FWT_Test.m:38-47:
    f 0.005 -> 0.03 cycles/sample over the central 80 %, Hann-tapered 40 % / 40 %."""
    n0, n1 = int(0.1 * nmax), int(0.9 * nmax)
    L = n1 - n0
    t = np.arange(L, dtype=np.float64)
    f0, f1 = 0.005, 0.03
    k = (f1 / f0) ** (1.0 / L)
    phase = 2 * np.pi * f0 * (k ** t - 1) / math.log(k)
    win = np.ones(L)
    e = int(0.4 * L)
    h = 0.5 * (1 - np.cos(np.pi * np.arange(e) / e))
    win[:e] = h
    win[L - e:] = h[::-1]
    s = np.zeros(nmax)
    s[n0:n1] = np.sin(phase) * win
    return s


def synthetic_traces(mtr: int, nmax: int, seed: int = 20171, amp: float = 0.5,
                     first: int = 0) -> np.ndarray:
    """x[k] = amp*s + N(0,1) noise keyed by (seed, global trace index k) -> [mtr, max] float32."""
    s = synthetic_signal(nmax) * amp
    out = np.empty((mtr, nmax), dtype=np.float32)
    for k in range(mtr):
        rng = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, 0, first + k]))
        out[k] = (s + rng.standard_normal(nmax)).astype(np.float32)
    return out


def nerr(a: np.ndarray, b: np.ndarray) -> float:
    """Normalised max error max|a-b| / max|b|."""
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
