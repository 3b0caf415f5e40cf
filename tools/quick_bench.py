"""Scratch timing of the device-resident plan API (not the official bench)."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import importlib.util
_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("ts_pws_b200", os.path.join(_root, "ts-pws_b200", "__init__.py"))
B = importlib.util.module_from_spec(_spec); sys.modules["ts_pws_b200"] = B; _spec.loader.exec_module(B)
lib = B.load_library(os.environ.get('TSPWS_LIB') or None)

def run(name, mtr, n, reps=5, **kw):
    p = B.default_params(**kw)
    lib.tspws_b200_derive_params(C.byref(p), n, 1.0)
    x = torch.randn(mtr, n, device="cuda", dtype=torch.float32)
    ls = torch.empty(n, device="cuda"); ts = torch.empty(n, device="cuda")
    plan = lib.tspws_b200_plan_create(C.byref(p), n, mtr, mtr, 0)
    assert plan
    side = torch.cuda.Stream(); torch.cuda.set_stream(side)
    stream = side.cuda_stream
    def step():
        rc = lib.tspws_b200_plan_stack(plan, x.data_ptr(), ls.data_ptr(), ts.data_ptr(), stream)
        assert rc == 0, rc
    for _ in range(2): step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        step(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts_ = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    best = min(ts_)
    if os.environ.get("PROF"):
        lib.tspws_b200_profile(1); lib.tspws_b200_profile_reset()
        for i in range(reps): step()
        torch.cuda.synchronize()
        prof = B.read_profile(lib); lib.tspws_b200_profile(0)
        print("   " + "  ".join(f"{k}={v[1]/max(v[0],1)*1e3:.1f}us(x{v[0]//reps})" for k, v in prof.items()))
    print(f"{name}: {mtr}x{n} J={p.J} V={p.V} best {best*1e3:.1f} us  median {sorted(ts_)[len(ts_)//2]*1e3:.1f} us  "
          f"{mtr*n/best/1e6:.3f} Gsamples/s  {mtr*n*4/best/1e6:.1f} GB/s", flush=True)
    lib.tspws_b200_plan_destroy(plan)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ts":
        run("two-stage morlet rm", 8192, 16384, reps=2, Kmax=10, unbiased=1, lrm=1)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ss64":
        run("single-stage morlet w0=6 (64 traces)", 64, 16384, reps=1, w0=6.0, w0set=4, lrm=1)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "c2":
        run("single-stage morlet w0=6 (8192 traces)", 8192, 16384, reps=3, w0=6.0, w0set=4, lrm=1)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ss":
        run("single-stage morlet w0=6 (256 traces)", 256, 16384, reps=1, w0=6.0, w0set=4, lrm=1)
        sys.exit(0)
    run("two-stage morlet rm", 8192, 16384, Kmax=10, unbiased=1, lrm=1)
    run("two-stage morlet", 8192, 16384, Kmax=10, unbiased=1)
    run("two-stage mexhat rm", 8192, 16384, type=-3, Kmax=10, unbiased=1, lrm=1)
    run("single-stage morlet w0=6 (512 traces)", 512, 16384, reps=3, w0=6.0, w0set=4, lrm=1)
