"""Worker for the multi-process tests (launched by torch.distributed.run).

mode "gloo-oracle": CPU, numpy oracle -- checks the sharding + exchange ALGEBRA the multi-GPU path relies on:
    every rank reduces its contiguous shard of traces to the exchange block (two-stage: group sums with the
    GLOBAL group index; single-stage: phasor / linear sums), the blocks are all-reduced (sum) and the tail
    runs replicated; the result must equal the one-process stack of all traces.
mode "nccl-plan": GPU, the real thing through the C-ABI plan API (one process per GPU).
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tspws_oracle as O  # noqa: E402


def shard(mtr, world, rank):
    per = mtr // world
    lo = rank * per
    hi = mtr if rank == world - 1 else lo + per
    return lo, hi


def gloo_oracle():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    mtr, n = 42, 512
    x = O.synthetic_traces(mtr, n, seed=77)
    lo, hi = shard(mtr, world, rank)
    for kw in (dict(Kmax=5, unbiased=1, lrm=1, fold=1), dict(lrm=1)):
        p = O.derive_params(O.Params(**kw), n, 1.0)
        fr = O.frame_from_params(p, n)
        full = O.tspws_main(O.Params(**kw), x)
        xs = x[lo:hi].copy()
        O.fold_inplace(xs) if p.fold else None
        O.remove_mean_inplace(xs) if p.lrm else None
        if p.Kmax:
            g = O.group_index(mtr, p.Kmax)[lo:hi]                 # GLOBAL trace index -> group
            block = np.zeros((p.Kmax, n))
            np.add.at(block, g, xs.astype(np.float64))
            t = torch.from_numpy(block)
            dist.all_reduce(t)
            st, ps = O.stacks(fr, t.numpy())
            K = p.Kmax
        else:
            st, ps = O.stacks(fr, xs)
            t = torch.from_numpy(np.stack([st.real, st.imag, ps.real, ps.imag]))
            dist.all_reduce(t)
            a = t.numpy()
            st, ps = a[0] + 1j * a[1], a[2] + 1j * a[3]
            K = mtr
        ts = O.synthesis(fr, O.weight(st, ps, K, mtr, p.wu, bool(p.unbiased))).astype(np.float32)
        ls = (O.synthesis(fr, st).astype(np.float32) / np.float32(mtr))
        assert O.nerr(ts, full["tsPWS"]) < 1e-6, (rank, kw, O.nerr(ts, full["tsPWS"]))
        assert O.nerr(ls, full["ls"]) < 1e-6
    dist.barrier()
    if rank == 0:
        print("GLOO_ORACLE_OK")
    dist.destroy_process_group()


def nccl_plan():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ts_pws_b200", os.path.join(ROOT, "ts-pws_b200", "__init__.py"))
    B = importlib.util.module_from_spec(spec)
    sys.modules["ts_pws_b200"] = B
    spec.loader.exec_module(B)
    lib = B.load_library()
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    lib.tspws_b200_set_option(B.OPT_DEVICE, local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    mtr, n = 256, 4096
    x = O.synthetic_traces(mtr, n, seed=91) + np.float32(0.1)
    lo, hi = shard(mtr, world, rank)
    stream = torch.cuda.current_stream().cuda_stream
    for kw in (dict(Kmax=10, unbiased=1, lrm=1, fold=1), dict(type=-3, Kmax=7, lrm=1), dict(lrm=1, w0=6.0, w0set=4)):
        p = B.default_params(**kw)
        lib.tspws_b200_derive_params(C.byref(p), n, 1.0)
        dx = torch.from_numpy(x[lo:hi].copy()).cuda()
        ls = torch.empty(n, device="cuda")
        ts = torch.empty(n, device="cuda")
        plan = lib.tspws_b200_plan_create(C.byref(p), n, hi - lo, mtr, lo)
        assert plan
        nbytes = lib.tspws_b200_plan_exchange_bytes(plan)
        is64 = lib.tspws_b200_plan_exchange_is_f64(plan)
        ex = torch.zeros(nbytes // (8 if is64 else 4), device="cuda", dtype=torch.float64 if is64 else torch.float32)
        assert lib.tspws_b200_plan_bind_exchange(plan, ex.data_ptr()) == 0
        for _ in range(2):                                           # twice: buffers must re-arm
            assert lib.tspws_b200_plan_reduce(plan, dx.data_ptr(), stream) == 0
            dist.all_reduce(ex)
            lib.tspws_b200_plan_mark_exchanged(plan)
            assert lib.tspws_b200_plan_finish(plan, ls.data_ptr(), ts.data_ptr(), stream) == 0
            torch.cuda.synchronize()
        kwo = {k: v for k, v in kw.items()}
        o = O.tspws_main(O.Params(**kwo), x)
        e_ts, e_ls = O.nerr(ts.cpu().numpy(), o["tsPWS"]), O.nerr(ls.cpu().numpy(), o["ls"])
        nc = lib.tspws_b200_plan_ncoef(plan)
        ps = np.empty(2 * nc)
        lib.tspws_b200_plan_get_stacks(plan, None, ps.ctypes.data_as(C.POINTER(C.c_double)), nc)
        e_c = O.nerr(np.hypot(ps[0::2], ps[1::2]) / lib.tspws_b200_plan_K(plan), o["coherence"])
        assert e_ts <= 1e-4 and e_ls <= 1e-4 and e_c <= 1e-5, (rank, kw, e_ts, e_ls, e_c)
        lib.tspws_b200_plan_destroy(plan)
        if not p.Kmax:
            continue
        # ---- the same stack with the peer exchange (CUDA IPC mappings, sum over ranks fused into the
        # first kernel of plan_finish) instead of the NCCL all-reduce: same result to fp64 rounding ----
        ref_ts, ref_ls = ts.clone(), ls.clone()
        plan = lib.tspws_b200_plan_create(C.byref(p), n, hi - lo, mtr, lo)
        h = (C.c_ubyte * 64)()
        assert lib.tspws_b200_plan_peer_export(plan, world, h) == 0
        mine = torch.tensor(list(h), dtype=torch.uint8, device="cuda")
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        buf = (C.c_ubyte * (64 * world))(*torch.cat(allh).cpu().tolist())
        assert lib.tspws_b200_plan_peer_import(plan, rank, world, buf) == 0
        side = torch.cuda.Stream()
        dist.barrier()
        for it in range(3):                                          # step parity, re-arming, graph replay
            assert lib.tspws_b200_plan_stack(plan, dx.data_ptr(), ls.data_ptr(), ts.data_ptr(), side.cuda_stream) == 0
            side.synchronize()
            assert lib.tspws_b200_plan_peer_error(plan) == 0
            d_ts, d_ls = O.nerr(ts.cpu().numpy(), ref_ts.cpu().numpy()), O.nerr(ls.cpu().numpy(), ref_ls.cpu().numpy())
            assert d_ts <= 1e-6 and d_ls <= 1e-6, (rank, kw, it, d_ts, d_ls)
        dist.barrier()                                               # nobody unmaps while a peer still reads
        assert lib.tspws_b200_plan_peer_close(plan) == 0
        dist.barrier()                                               # nobody frees while a peer still maps
        lib.tspws_b200_plan_destroy(plan)
    dist.barrier()
    if rank == 0:
        print("NCCL_PLAN_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    {"gloo-oracle": gloo_oracle, "nccl-plan": nccl_plan}[sys.argv[1]]()
