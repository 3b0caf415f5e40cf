"""Two-stream software pipeline over independent headline stacks: all group-sum kernels back to back on stream H,
the tails (plan_finish) on stream T, NP plans as buffers (scratch)."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, importlib.util
_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("ts_pws_b200", os.path.join(_root, "ts-pws_b200", "__init__.py"))
B = importlib.util.module_from_spec(_spec); sys.modules["ts_pws_b200"] = B; _spec.loader.exec_module(B)
lib = B.load_library()
mtr, n = 8192, 16384
def run(NP, reps=60, hiprio=True, **kw):
    p = B.default_params(**kw)
    lib.tspws_b200_derive_params(C.byref(p), n, 1.0)
    xs = [torch.randn(mtr, n, device="cuda") for _ in range(2)]
    outs = [(torch.empty(n, device="cuda"), torch.empty(n, device="cuda")) for _ in range(NP)]
    plans = [lib.tspws_b200_plan_create(C.byref(p), n, mtr, mtr, 0) for _ in range(NP)]
    H = torch.cuda.Stream(); T = torch.cuda.Stream(priority=-1 if hiprio else 0)
    reduced = [torch.cuda.Event() for _ in range(NP)]; finished = [torch.cuda.Event() for _ in range(NP)]
    def submit(i, first):
        k = i % NP
        if not first: H.wait_event(finished[k])
        assert lib.tspws_b200_plan_reduce(plans[k], xs[i & 1].data_ptr(), H.cuda_stream) == 0
        reduced[k].record(H)
        T.wait_event(reduced[k])
        assert lib.tspws_b200_plan_finish(plans[k], outs[k][0].data_ptr(), outs[k][1].data_ptr(), T.cuda_stream) == 0
        finished[k].record(T)
    for i in range(3 * NP): submit(i, i < NP)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        start = torch.cuda.Event(enable_timing=True); end = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record(H)
        t0 = time.perf_counter()
        for i in range(reps): submit(i, False)
        host = (time.perf_counter() - t0) / reps
        H.wait_event(finished[(reps - 1) % NP]); 
        for k in range(NP): H.wait_event(finished[k])
        end.record(H)
        torch.cuda.synchronize()
        best = min(best, start.elapsed_time(end) / reps)
    print(f"pipeline depth {NP} prio {hiprio}: {best*1e3:.1f} us per stack, {mtr*n*4/best/1e6:.0f} GB/s  (host issue {host*1e6:.1f} us per stack)", flush=True)
    for pl in plans: lib.tspws_b200_plan_destroy(pl)
for NP in (1, 2, 3):
    run(NP, Kmax=10, unbiased=1, lrm=1)
run(2, hiprio=False, Kmax=10, unbiased=1, lrm=1)
