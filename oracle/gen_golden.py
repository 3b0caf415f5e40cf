"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref.

Run in the build container (where /root/reference exists):
    make -C oracle && python -m oracle.gen_golden
For every case of oracle/cases.py it stores what the reference's tspws_main returns (ls, tsPWS,
the derived parameters) and the ST / PS stacks taken from the reference's link-visible
tspws_stacks_float / partial_linear_stacks + tspws_stacks_double, plus a checksum of the input.
TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import hashlib
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases as CS, ref_lib as R  # noqa: E402


def _pkg():
    spec = importlib.util.spec_from_file_location("ts_pws_b200", os.path.join(ROOT, "ts-pws_b200", "__init__.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["ts_pws_b200"] = m
    spec.loader.exec_module(m)
    return m


def run_reference(B, ref, case):
    x = CS.make_input(case)
    p = B.default_params(**case["params"])
    r = B.call_tspws_main(ref, x, p, dt=case["dt"], beg=case.get("beg"))
    assert r["rc"] == 0
    pp = r["params"]
    mtr = pp.Nmax if pp.Nmax else case["mtr"]
    rf = R.RefFrame(pp.type, pp.J, pp.V, case["n"], pp.s0, pp.b0, pp.w0, pp.uni)
    st, ps, K = rf.stacks(r["sigall"][:mtr], pp.Kmax)
    rf.close()
    derived = np.array([pp.J, pp.V, pp.s0, pp.b0, pp.w0, pp.fmin, pp.fold], dtype=np.float64)
    return dict(ls=r["ls"], tsPWS=r["tsPWS"], st=st, ps=ps, K=np.int64(K), derived=derived,
                input_sha=np.frombuffer(hashlib.sha256(x.tobytes()).digest(), dtype=np.uint8))


def run_reference_resample(B, ref, case):
    """Reference resampling outputs + the selection rows it used (single-threaded: its two-stage
    sampler calls rand() inside an OpenMP loop, src/ts_pws1f_lib.c:640)."""
    import ctypes as C
    import math
    C.CDLL("libgomp.so.1").omp_set_num_threads(1)
    libc = C.CDLL(None)
    x = CS.make_input(case)
    p = B.default_params(**case["params"])
    mtr = case["mtr"]
    times = None
    if case["plans"][0] == "rand":
        libc.srand(case["plans"][1])
        r = B.call_tspws_main(ref, x, p, dt=case["dt"])
        libc.srand(case["plans"][1])
        sel = R.subsampling_plans(p.subsmpl_N, mtr, int(math.ceil(mtr * p.subsmpl_p)))
    else:
        times = CS.jackknife_times(mtr)
        r = B.call_tspws_main(ref, x, p, dt=case["dt"], times=times)
        rc, sel = R.jackknife_plans(times, p.jackknife_n, p.jackknife_d, math.comb(p.jackknife_n, p.jackknife_d))
        assert rc == 0
    assert r["rc"] == 0
    return dict(ls=r["ls"], tsPWS=r["tsPWS"], ls_subsmpl=r["ls_subsmpl"], tsPWS_subsmpl=r["tsPWS_subsmpl"],
                mtr_subsmpl=r["mtr_subsmpl"], sel=sel.astype(np.int8),
                input_sha=np.frombuffer(hashlib.sha256(x.tobytes()).digest(), dtype=np.uint8))


def run_reference_convergence(B, ref, case):
    x = CS.make_input(case)
    p = B.default_params(**case["params"])
    r = B.call_tspws_main(ref, x, p, dt=case["dt"])
    assert r["rc"] == 0
    g = dict(ls=r["ls"], tsPWS=r["tsPWS"], input_sha=np.frombuffer(hashlib.sha256(x.tobytes()).digest(), dtype=np.uint8))
    for k in ("ls_sim", "tsPWS_sim", "ls_misfit", "tsPWS_misfit", "ls_steps", "tsPWS_steps"):
        if k in r:
            g[k] = r[k]
    return g


REAL_DIR = "/root/reference/examples/ECH.00Z.CAN.00Z_500days"
# the three runs of the reference's examples/example.sh:16,27,39 on the first 60 daily correlations
REAL_CASES = {
    "ex1_default": dict(),
    "ex2_fmin_J3": dict(wu=2.0, lrm=1, fold=1, fmin=0.004, J=3),
    "ex3_two_stage_unbiased": dict(wu=2.0, lrm=1, fold=1, fmin=0.004, J=3, Kmax=10, unbiased=1),
}


def read_sac_py(path):
    """Independent little-endian SAC reader (struct-level) used to cross-check the C loader."""
    raw = open(path, "rb").read()
    f = np.frombuffer(raw[:280], dtype="<f4")
    i = np.frombuffer(raw[280:440], dtype="<i4")
    npts = int(i[9])
    return dict(delta=float(f[0]), b=float(f[5]), npts=npts, data=np.frombuffer(raw[632:632 + 4 * npts], dtype="<f4").copy())


def run_reference_real(B, ref):
    files = sorted(os.listdir(REAL_DIR))[:60]
    traces = [read_sac_py(os.path.join(REAL_DIR, f)) for f in files]
    x = np.stack([t["data"] for t in traces]).astype(np.float32)
    dt, beg = np.float32(traces[0]["delta"]), np.float32(traces[0]["b"])
    g = dict(x=x, dt=dt, beg=beg)
    for name, kw in REAL_CASES.items():
        r = B.call_tspws_main(ref, x, B.default_params(**kw), dt=float(dt), beg=float(beg))
        assert r["rc"] == 0
        pp = r["params"]
        rf = R.RefFrame(pp.type, pp.J, pp.V, x.shape[1], pp.s0, pp.b0, pp.w0, pp.uni)
        st, ps, K = rf.stacks(r["sigall"], pp.Kmax)
        rf.close()
        g[name + "_ls"], g[name + "_tsPWS"] = r["ls"], r["tsPWS"]
        g[name + "_ps"], g[name + "_st"], g[name + "_K"] = ps.astype(np.complex128), st.astype(np.complex128), np.int64(K)
        g[name + "_derived"] = np.array([pp.J, pp.V, pp.s0, pp.b0, pp.w0, pp.fmin, pp.fold], dtype=np.float64)
    return g


def main():
    B = _pkg()
    ref = B.declare_reference(R.load())
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name, case in CS.RESAMPLE_CASES.items():
        g = run_reference_resample(B, ref, case)
        np.savez_compressed(os.path.join(out, name + ".npz"), **g)
        print(f"{name}: M={g['sel'].shape[0]} K={g['sel'].sum(axis=1).tolist()} max|ts_sub|={np.abs(g['tsPWS_subsmpl']).max():.3e}")
    if os.path.isdir(REAL_DIR):
        g = run_reference_real(B, ref)
        np.savez_compressed(os.path.join(out, "real_ech_can_60.npz"), **g)
        print("real_ech_can_60:", g["x"].shape, "dt", g["dt"], "beg", g["beg"], [k for k in g if k.endswith("_K")])
    for name, case in CS.CONVERGENCE_CASES.items():
        g = run_reference_convergence(B, ref, case)
        np.savez_compressed(os.path.join(out, name + ".npz"), **g)
        print(f"{name}: sim[-1]={g['tsPWS_sim'][-1]:.6f} misfit[0]={g['tsPWS_misfit'][0]:.3e}")
    for name, case in CS.CASES.items():
        g = run_reference(B, ref, case)
        # ST/PS are stored as complex64-pairs of float64? keep float64 for the 1e-5 coherence check,
        # but drop to float32 where ncoef is large to keep the fixtures small
        np.savez_compressed(os.path.join(out, name + ".npz"), **g)
        print(f"{name}: ncoef={g['st'].size} K={int(g['K'])} max|tsPWS|={np.abs(g['tsPWS']).max():.3e}")


if __name__ == "__main__":
    main()
