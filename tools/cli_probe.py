"""Runs tspws_b200_cli with the resampling / convergence flags on a small .bin file and checks the files it
writes (names and sizes as in the reference CLI, src/ts_pws1f.c:354-418) against tspws_main called directly."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import cases as CS, tspws_oracle as O   # noqa: E402
import conftest                                      # noqa: E402

B = conftest.load_pkg()
lib = B.load_library()
d = tempfile.mkdtemp()
mtr, n = 40, 512
x = O.synthetic_traces(mtr, n, seed=3)
times = CS.jackknife_times(mtr)
dat = B.Data()
dat.sigall = x.ctypes.data_as(C.POINTER(C.c_float))
dat.hdr.max, dat.hdr.mtr, dat.hdr.dt, dat.hdr.beg = n, mtr, 1.0, -(n - 1) / 2.0
t = np.ascontiguousarray(times, dtype=np.int64)
dat.time = t.ctypes.data_as(C.POINTER(C.c_long))
binf = os.path.join(d, "a.bin")
assert lib.tspws_b200_save_bin(binf.encode(), C.byref(dat), b"pcc") == 0
cli = os.path.join(ROOT, "ts-pws_b200", "csrc", "tspws_b200_cli")


def run(*args):
    r = subprocess.run([cli, binf, "bin", *args], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, (args, r.stdout[-500:], r.stderr[-500:])


def sac_data(path, npts):
    raw = open(path, "rb").read()
    assert len(raw) == 632 + 4 * npts, (path, len(raw))
    return np.frombuffer(raw[632:], dtype=np.float32)


run("rm", "osac=c1", "convergence", "AllSteps")
r = B.call_tspws_main(lib, x, B.default_params(lrm=1, convergence=1, AllSteps=1))
for name, key, dt in (("ts_pws_c1_convergence", "tsPWS_sim", np.float64), ("tl_c1_misfit", "ls_misfit", np.float64),
                      ("ts_pws_c1_steps", "tsPWS_steps", np.float32)):
    a = np.fromfile(os.path.join(d, name), dtype=dt)
    print(name, a.shape, "max diff vs tspws_main", float(np.max(np.abs(a - r[key].ravel()))))
print("ts_pws_c1.sac vs tspws_main", float(np.max(np.abs(sac_data(os.path.join(d, "ts_pws_c1.sac"), n) - r["tsPWS"]))))

run("rm", "osac=j1", "TwoStage=4", "unbiased", "jackknife_n=6", "jackknife_d=2", "obin")
r = B.call_tspws_main(lib, x, B.default_params(lrm=1, Kmax=4, unbiased=1, jackknife_n=6, jackknife_d=2), times=times)
raw = open(os.path.join(d, "ts_pws_j1_subsmpl.bin"), "rb").read()
M = 15
assert len(raw) == 116 + M * 8 + M * n * 4, len(raw)
print("jackknife obin: K rows", np.array_equal(np.frombuffer(raw[116:116 + 8 * M], dtype=np.int64), r["mtr_subsmpl"]),
      "max diff", float(np.max(np.abs(np.frombuffer(raw[116 + 8 * M:], dtype=np.float32).reshape(M, n) - r["tsPWS_subsmpl"]))))

run("rm", "fold", "osac=s1", "subsmpl_N=3", "subsmpl_prob=0.5")
for m in range(3):
    a = sac_data(os.path.join(d, f"ts_pws_s1_subsmpl_{m}.sac"), (n + 1) // 2)
    b = sac_data(os.path.join(d, f"tl_s1_subsmpl_{m}.sac"), (n + 1) // 2)
    assert np.isfinite(a).all() and np.isfinite(b).all() and a.any() and b.any()
print("sub-sampling SAC rows written: ok")
print("CLI_PROBE_OK")
