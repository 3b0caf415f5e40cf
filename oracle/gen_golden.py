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


def main():
    B = _pkg()
    ref = B.declare_reference(R.load())
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name, case in CS.CASES.items():
        g = run_reference(B, ref, case)
        # ST/PS are stored as complex64-pairs of float64? keep float64 for the 1e-5 coherence check,
        # but drop to float32 where ncoef is large to keep the fixtures small
        np.savez_compressed(os.path.join(out, name + ".npz"), **g)
        print(f"{name}: ncoef={g['st'].size} K={int(g['K'])} max|tsPWS|={np.abs(g['tsPWS']).max():.3e}")


if __name__ == "__main__":
    main()
