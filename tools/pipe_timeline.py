"""Per-CTA timeline of the stack pipeline (timeline build of the library in tools/_tl): who runs where, when."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, importlib.util, numpy as np
_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["TSPWS_TIMELINE"] = "1"
_spec = importlib.util.spec_from_file_location("ts_pws_b200", os.path.join(_root, "ts-pws_b200", "__init__.py"))
B = importlib.util.module_from_spec(_spec); sys.modules["ts_pws_b200"] = B; _spec.loader.exec_module(B)
lib = B.load_library(os.path.join(_root, "tools", "_tl", "libtspws_b200.so"))
lib.tspws_cu_timeline_flush.argtypes = [C.c_char_p]
mtr, n, depth, nst = 8192, 16384, int(os.environ.get("DEPTH", "4")), 24
kw = dict(Kmax=10, unbiased=1, lrm=1)
if os.environ.get("MEXHAT"): kw["type"] = -3
p = B.default_params(**kw)
lib.tspws_b200_derive_params(C.byref(p), n, 1.0)
xs = [torch.randn(mtr, n, device="cuda") for _ in range(2)]
outs = [(torch.empty(n, device="cuda"), torch.empty(n, device="cuda")) for _ in range(depth)]
pipe = lib.tspws_b200_pipe_create(C.byref(p), n, mtr, mtr, 0, depth)
def submit(i):
    k = i % depth
    assert lib.tspws_b200_pipe_submit(pipe, xs[i & 1].data_ptr(), outs[k][0].data_ptr(), outs[k][1].data_ptr()) == 0
for i in range(3 * depth): submit(i)
lib.tspws_b200_pipe_sync(pipe)
out = os.environ.get("OUT", "gpurun_out/pipe_timeline.txt")
lib.tspws_cu_timeline_flush(b"/dev/null")
for i in range(nst): submit(i)
lib.tspws_b200_pipe_sync(pipe)
lib.tspws_cu_timeline_flush(out.encode())
rec = np.loadtxt(out, dtype=np.int64)            # kid cta start end sm
t0 = rec[:, 2].min()
names = {0: "colsum2", 1: "post2", 2: "fwd_tail", 3: "combine", 4: "inv2", 5: "inv4", 6: "finish"}
print("records", len(rec), "span us", (rec[:, 3].max() - t0) / 1e3, "per stack", (rec[:, 3].max() - t0) / 1e3 / nst)
for kid in sorted(names):
    r = rec[rec[:, 0] == kid]
    if not len(r): continue
    starts = np.sort(r[r[:, 1] == 0][:, 2])
    idx = np.searchsorted(starts, r[:, 2] + 3000, side="right") - 1
    rows = []
    for L in range(len(starts)):
        q = r[idx == L]
        if not len(q): continue
        rows.append(((q[:, 2].min() - t0) / 1e3, (q[:, 3].max() - t0) / 1e3, len(q), np.mean(q[:, 3] - q[:, 2]) / 1e3,
                     (np.percentile(q[:, 2], 50) - t0) / 1e3))
    print(names[kid], "launches", len(rows))
    for a, b, c, m, med in rows[6:12]:
        print(f"    start {a:8.1f} end {b:8.1f} dur {b - a:6.1f} us  ctas {c:4d}  mean cta life {m:6.2f} us  median cta start {med:8.1f}")
