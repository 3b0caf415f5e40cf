"""Per-CTA timeline of the stack pipeline over the peer exchange (run under torchrun, timeline build in tools/_tl):
   TSPWS_LIB=tools/_tl/libtspws_b200.so TSPWS_TIMELINE=1 python -m torch.distributed.run --nproc-per-node 2 ... tools/pipe_timeline_mg.py"""
import ctypes as C, os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
args = types.SimpleNamespace(gpus=int(os.environ.get("WORLD_SIZE", "1")))
ctx = bench.Ctx(args)
torch, lib = ctx.torch, ctx.lib
lib.tspws_cu_timeline_flush.argtypes = [C.c_char_p]
mtr, n, depth, nst = 8192, 16384, int(os.environ.get("DEPTH", "4")), 24
wl = bench.WORKLOADS["two_stage_morlet"]
x = [torch.randn(mtr, n, device="cuda") for _ in range(2)]
ps = bench.PipedStacks(ctx, wl["params"], n, mtr, mtr * ctx.world, ctx.rank * mtr, depth=depth)
assert ps.ok, "pipe set-up failed"
for i in range(3 * depth): ps.submit(x[i & 1])
ps.sync(); ctx.barrier()
lib.tspws_cu_timeline_flush(b"/dev/null")
ms = ps.timed(x, nst)
out = f"gpurun_out/tl_mg_rank{ctx.rank}.txt"
lib.tspws_cu_timeline_flush(out.encode())
if ctx.rank == 0:
    print("ms per stack", ms, "peer", ps.peer, flush=True)
ps.close(); ctx.shutdown()
