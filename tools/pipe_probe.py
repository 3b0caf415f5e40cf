"""Throughput of a stream of independent headline stacks with NP plans in flight on NP streams (scratch)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, importlib.util
_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("ts_pws_b200", os.path.join(_root, "ts-pws_b200", "__init__.py"))
B = importlib.util.module_from_spec(_spec); sys.modules["ts_pws_b200"] = B; _spec.loader.exec_module(B)
lib = B.load_library()
mtr, n = 8192, 16384
def run(NP, reps=40, prio=False, **kw):
    p = B.default_params(**kw)
    lib.tspws_b200_derive_params(C.byref(p), n, 1.0)
    xs = [torch.randn(mtr, n, device="cuda") for _ in range(NP)]
    outs = [(torch.empty(n, device="cuda"), torch.empty(n, device="cuda")) for _ in range(NP)]
    plans = [lib.tspws_b200_plan_create(C.byref(p), n, mtr, mtr, 0) for _ in range(NP)]
    streams = [torch.cuda.Stream() for _ in range(NP)]
    def step(i):
        k = i % NP
        rc = lib.tspws_b200_plan_stack(plans[k], xs[k].data_ptr(), outs[k][0].data_ptr(), outs[k][1].data_ptr(), streams[k].cuda_stream)
        assert rc == 0
    for i in range(4 * NP): step(i)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        start = torch.cuda.Event(enable_timing=True); ends = [torch.cuda.Event(enable_timing=True) for _ in range(NP)]
        torch.cuda.synchronize()
        start.record(streams[0])
        for s in streams[1:]: s.wait_event(start)
        for i in range(reps): step(i)
        for k in range(NP): ends[k].record(streams[k])
        torch.cuda.synchronize()
        t = max(start.elapsed_time(e) for e in ends) / reps
        best = min(best, t)
    print(f"{NP} stacks in flight: {best*1e3:.1f} us per stack, {mtr*n*4/best/1e6:.0f} GB/s", flush=True)
    for pl in plans: lib.tspws_b200_plan_destroy(pl)
for NP in (1, 2, 3):
    run(NP, Kmax=10, unbiased=1, lrm=1)
