"""BASELINE configs[4]: independent station pairs of 500 traces x 8192 samples, default
Morlet (40 scales), nu = 2, from PINNED host memory: tspws_b200_main_batch (pipelined) against a loop
of tspws_main calls.  Prints pairs/s and trace-samples/s.     python tools/batch_bench.py [npairs] [Kmax]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest                                      # noqa: E402

B = conftest.load_pkg()
lib = B.load_library()
npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kmax = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mtr, n = 500, 8192
g = torch.Generator().manual_seed(5)
host = torch.randn(npairs, mtr, n, generator=g).pin_memory()
ls = torch.zeros(npairs, n).pin_memory()
ts = torch.zeros(npairs, n).pin_memory()
xs = [host[i].numpy() for i in range(npairs)]
lsn = [ls[i].numpy() for i in range(npairs)]
tsn = [ts[i].numpy() for i in range(npairs)]
kw = dict(lrm=1, Kmax=kmax, unbiased=1 if kmax else 0)
p = B.default_params(**kw)
for rep in range(3):
    t0 = time.perf_counter()
    r = B.call_main_batch(lib, xs, p, out_ls=lsn, out_ts=tsn)
    t1 = time.perf_counter()
    assert r["rc"] == 0
    tb = t1 - t0
ref_ts = ts.clone()
outs, dats = [], []
for i in range(npairs):
    out, dat = B.TsPWSOut(), B.Data()
    dat.sigall = C.cast(host[i].data_ptr(), C.POINTER(C.c_float))
    dat.hdr.max, dat.hdr.mtr, dat.hdr.dt, dat.hdr.beg = n, mtr, 1.0, -(n - 1) / 2.0
    out.ls = C.cast(ls[i].data_ptr(), C.POINTER(C.c_float))
    out.tsPWS = C.cast(ts[i].data_ptr(), C.POINTER(C.c_float))
    out.N, out.mtr = n, mtr
    outs.append(out); dats.append(dat)
for rep in range(3):
    t0 = time.perf_counter()
    for i in range(npairs):
        pp = B.TsPWS.from_buffer_copy(p)
        assert lib.tspws_main(C.byref(pp), C.byref(outs[i]), C.byref(dats[i])) == 0
    t1 = time.perf_counter()
    tl = t1 - t0
assert torch.equal(ref_ts, ts)
print(f"{npairs} pairs x {mtr} x {n} Kmax={kmax}: batch {tb*1e3:.1f} ms = {npairs/tb:.0f} pairs/s = {npairs*mtr*n/tb:.3e} samples/s"
      f" ({npairs*mtr*n*4/tb/1e9:.1f} GB/s H2D);  loop {tl*1e3:.1f} ms = {npairs/tl:.0f} pairs/s = {npairs*mtr*n/tl:.3e} samples/s")
