"""tspws_b200_pipe_* throughput over independent headline stacks (scratch)."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, importlib.util
_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("ts_pws_b200", os.path.join(_root, "ts-pws_b200", "__init__.py"))
B = importlib.util.module_from_spec(_spec); sys.modules["ts_pws_b200"] = B; _spec.loader.exec_module(B)
lib = B.load_library(os.environ.get('TSPWS_LIB') or None)
mtr, n = 8192, 16384
def run(depth, reps=int(os.environ.get('REPS', '60')), **kw):
    p = B.default_params(**kw)
    lib.tspws_b200_derive_params(C.byref(p), n, 1.0)
    xs = [torch.randn(mtr, n, device="cuda") for _ in range(2)]
    outs = [(torch.empty(n, device="cuda"), torch.empty(n, device="cuda")) for _ in range(depth)]
    pipe = lib.tspws_b200_pipe_create(C.byref(p), n, mtr, mtr, 0, depth)
    assert pipe
    H = torch.cuda.ExternalStream(lib.tspws_b200_pipe_stream(pipe, 0)); T = torch.cuda.ExternalStream(lib.tspws_b200_pipe_stream(pipe, 1))
    def submit(i):
        k = i % depth
        assert lib.tspws_b200_pipe_submit(pipe, xs[k & 1].data_ptr(), outs[k][0].data_ptr(), outs[k][1].data_ptr()) == 0
    for i in range(3 * depth): submit(i)
    assert lib.tspws_b200_pipe_sync(pipe) == 0
    ts = []
    for rep in range(5):
        start = torch.cuda.Event(enable_timing=True); end = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record(H)
        t0 = time.perf_counter()
        for i in range(reps): submit(i)
        host = (time.perf_counter() - t0) / reps
        end.record(T)
        lib.tspws_b200_pipe_sync(pipe); torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / reps * 1e3
        ts.append(max(start.elapsed_time(end) / reps, wall if os.environ.get('TSPWS_PIPE_TAILS') == '2' else 0.0))
    best = min(ts)
    print(f"pipe depth {depth}: best {best*1e3:.1f} median {sorted(ts)[2]*1e3:.1f} us per stack, {mtr*n*4/best/1e6:.0f} GB/s  (host issue {host*1e6:.1f} us per stack)", flush=True)
    lib.tspws_b200_pipe_destroy(pipe)
kw = dict(Kmax=10, unbiased=1, lrm=1)
if os.environ.get('MEXHAT'): kw['type'] = -3
for d in [int(v) for v in os.environ.get('DEPTHS', '1,2,3,4').split(',')]:
    run(d, **kw)
